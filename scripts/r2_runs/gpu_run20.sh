#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_lifecycle.py -m gpu -q -x > gpurun_out/r2_pytest20.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest20.log
tail -15 gpurun_out/r2_pytest20.log
