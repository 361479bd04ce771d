#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest7.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest7.log
tail -4 gpurun_out/r2_pytest7.log
: > gpurun_out/r2_exp7.jsonl
for wl in c3 c1 c2 c4; do timeout 300 python scripts/exp.py $wl split >> gpurun_out/r2_exp7.jsonl 2>> gpurun_out/r2_exp7.err; done
cat gpurun_out/r2_exp7.jsonl; tail -3 gpurun_out/r2_exp7.err
