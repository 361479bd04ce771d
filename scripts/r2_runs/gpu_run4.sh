#!/bin/bash
# 1-GPU: the gpu tests touched by the primary-ray cull (+ all renders), then A/B of the cull and of frames per batch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest4.log
tail -4 gpurun_out/r2_pytest4.log
: > gpurun_out/r2_exp4.jsonl
for wl in c3 c1 c4; do
  RBRT_NO_PRIMARY_CULL=1 timeout 300 python scripts/exp.py $wl nocull >> gpurun_out/r2_exp4.jsonl 2>> gpurun_out/r2_exp4.err
  timeout 300 python scripts/exp.py $wl cull >> gpurun_out/r2_exp4.jsonl 2>> gpurun_out/r2_exp4.err
done
cat gpurun_out/r2_exp4.jsonl
for fpb in 1 2 4; do FPB=$fpb DEPTHS=2 timeout 300 python scripts/pipe_time.py c3 1 8 2>&1 | tail -1; done
