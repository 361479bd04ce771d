#!/bin/bash
: > gpurun_out/r2_exp16.jsonl
timeout 300 python scripts/exp.py c3 base >> gpurun_out/r2_exp16.jsonl 2>> gpurun_out/r2_exp16.err
for v in v12 v23 v32 fs; do RBRT_GPU_LIB=$PWD/rbrt_b200/variants/librbrt_gpu_$v.so timeout 300 python scripts/exp.py c3 $v >> gpurun_out/r2_exp16.jsonl 2>> gpurun_out/r2_exp16.err; done
for t in 4 6 12; do RBRT_FETCH_THRESHOLD=$t timeout 300 python scripts/exp.py c3 thr$t >> gpurun_out/r2_exp16.jsonl 2>> gpurun_out/r2_exp16.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r2_exp16.jsonl'):
    d=json.loads(l); print(d['label'], 'frame', d['ms_frame_1'], 'trace', d['ms_trace_1'], 'split', d['ms_split_1'], 'frame8', d['ms_frame_8'], 'trace8', d['ms_trace_8'], d['checksum'])
PY
