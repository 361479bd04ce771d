#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
