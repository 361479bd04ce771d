#!/bin/bash
# L2 cache-hint experiment of round 2 (needed the RBRT_HINTS bit-mask macro of that commit; only the node hint was kept, as RBRT_NODE_L2_EVICT_LAST): C3 through scripts/exp.py, one line per variant.
mkdir -p gpurun_out; : > gpurun_out/r2_exp_l2_hints.jsonl
python scripts/exp.py c3 base >> gpurun_out/r2_exp_l2_hints.jsonl 2>> gpurun_out/r2_exp_l2_hints.err
for h in ${HINTS:-1 5 13 15 7}; do
  RBRT_GPU_LIB=$PWD/rbrt_b200/variants/librbrt_gpu_h$h.so python scripts/exp.py c3 hints$h >> gpurun_out/r2_exp_l2_hints.jsonl 2>> gpurun_out/r2_exp_l2_hints.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2_exp_l2_hints.jsonl'):
    d=json.loads(l); print(d['label'], 'frame', d['ms_frame_1'], 'trace', d['ms_trace_1'], 'split', d['ms_split_1'], '| 1/8:', d['ms_frame_8'], d['ms_trace_8'], d['ms_split_8'], 'checksum', d['checksum'])
PY
