#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest21.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest21.log
tail -12 gpurun_out/r2_pytest21.log
: > gpurun_out/r2_exp21.jsonl
for wl in c4 c3 c1; do timeout 300 python scripts/exp.py $wl groups >> gpurun_out/r2_exp21.jsonl 2>> gpurun_out/r2_exp21.err; done
python - <<'PY'
import json
for l in open('gpurun_out/r2_exp21.jsonl'):
    d=json.loads(l); print(d['wl'], 'frame', d['ms_frame_1'], 'trace', d['ms_trace_1'], 'split', d['ms_split_1'], 'frame8', d['ms_frame_8'], d['checksum'])
PY
