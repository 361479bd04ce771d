#!/bin/bash
# A/B of a run-time knob: scripts/exp.py per workload in $WLS with and without the environment assignment $1 (e.g. RBRT_NO_PRIMARY_CULL=1), twice
mkdir -p gpurun_out; out=gpurun_out/r2_exp_env_ab.jsonl; : > $out
for rep in 1 2; do for wl in ${WLS:-c3}; do
  python scripts/exp.py $wl default | grep "^{" >> $out 2>> gpurun_out/r2_exp_env_ab.err
  env "$1" python scripts/exp.py $wl "$1" | grep "^{" >> $out 2>> gpurun_out/r2_exp_env_ab.err
done; done
python - $out <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d['wl'], d['label'], 'frame', d['ms_frame_1'], 'trace', d['ms_trace_1'], 'split', d['ms_split_1'], '| 1/8:', d['ms_frame_8'], d['ms_trace_8'], d['ms_split_8'], 'checksum', d['checksum'])
PY
