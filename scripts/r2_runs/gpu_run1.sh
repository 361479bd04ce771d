#!/bin/bash
# GPU run 1 of round 2: the whole -m gpu suite, then the SAH-pass sweep on C3 / C2 / C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
: > gpurun_out/r2_exp1.jsonl
for wl in c3 c2 c4; do
  NO_SAH=1 timeout 300 python scripts/exp.py $wl nosah >> gpurun_out/r2_exp1.jsonl 2>> gpurun_out/r2_exp1.err
  for p in 1 2 3; do
    RBRT_SAH_PASSES=$p timeout 300 python scripts/exp.py $wl sah$p >> gpurun_out/r2_exp1.jsonl 2>> gpurun_out/r2_exp1.err
  done
done
cat gpurun_out/r2_exp1.jsonl
