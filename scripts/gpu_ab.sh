#!/bin/bash
# A/B of one experiment build (rbrt_b200/variants/librbrt_gpu_$1.so) against the in-tree library: scripts/exp.py on the workloads in $WLS, twice each
mkdir -p gpurun_out; out=gpurun_out/r2_exp_ab_$1.jsonl; : > $out
for rep in 1 2; do for wl in ${WLS:-c3}; do
  python scripts/exp.py $wl tree | grep "^{" >> $out 2>> gpurun_out/r2_exp_ab.err
  RBRT_GPU_LIB=$PWD/rbrt_b200/variants/librbrt_gpu_$1.so python scripts/exp.py $wl $1 | grep "^{" >> $out 2>> gpurun_out/r2_exp_ab.err
done; done
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d['wl'], d['label'], 'frame', d['ms_frame_1'], 'trace', d['ms_trace_1'], 'split', d['ms_split_1'], '| 1/8:', d['ms_frame_8'], d['ms_trace_8'], d['ms_split_8'], 'checksum', d['checksum'])
PY
