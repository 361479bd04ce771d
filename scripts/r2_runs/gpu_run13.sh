#!/bin/bash
mkdir -p gpurun_out
# builder edge cases + the big-mesh tests first (the collapse changed), then the SAH threshold sweep
timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/r2_pytest13.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest13.log
tail -3 gpurun_out/r2_pytest13.log
: > gpurun_out/r2_exp13.jsonl
for m in 1 4 16 64 256; do RBRT_SAH_MIN_LEAVES=$m timeout 300 python scripts/exp.py c3 sahmin$m >> gpurun_out/r2_exp13.jsonl 2>> gpurun_out/r2_exp13.err; done
NO_SAH=1 timeout 300 python scripts/exp.py c3 nosah >> gpurun_out/r2_exp13.jsonl 2>> gpurun_out/r2_exp13.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_exp13.jsonl'):
    d=json.loads(l); print(d['label'], 'build', d['build_ms'], 'V', d['V'], 'T', d['T'], 'frame', d['ms_frame_1'], 'trace', d['ms_trace_1'], 'nodes', d['nodes'], d['checksum'])
PY
DEPTH=2 timeout 200 python scripts/e2e_shard_time.py c3 8 12 2>&1 | tail -1
NO_SAH=1 DEPTH=2 timeout 200 python scripts/e2e_shard_time.py c3 8 12 2>&1 | tail -1
