#!/bin/bash
mkdir -p gpurun_out
for lib in main nopf; do
  if [ $lib = nopf ]; then export RBRT_GPU_LIB=$PWD/rbrt_b200/variants/librbrt_gpu_nopf.so; fi
  for n in 8 1; do
    echo "== lib $lib shard 1/$n"; timeout 300 python scripts/shard_iters.py $n 2>&1 | grep -E "tail kernel ran|TOTAL|it  [0-9] |it 1[01] " | tail -16
  done
done
