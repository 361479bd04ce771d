#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest6.log
tail -4 gpurun_out/r2_pytest6.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_n1_b.json 2> gpurun_out/r2_bench_c3_n1_b.err; echo "bench exit $?"
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2_bench_c3_n1_b.json'))
print('value',round(b['value']),'ms/step',round(b['ms_per_step'],3),'single',round(b['single_frame']['ms_per_step'],3),'fpb1',round(b['frames_per_batch_1']['ms_per_step'],3),'e2e',round(b['e2e']['value']),round(b['e2e']['ms_per_step'],3),b['image_check']['bit_identical'],b['cpu_baseline'])
PY
tail -2 gpurun_out/r2_bench_c3_n1_b.err
: > gpurun_out/r2_exp6.jsonl
for wl in c1 c3; do timeout 300 python scripts/exp.py $wl async >> gpurun_out/r2_exp6.jsonl 2>> gpurun_out/r2_exp6.err; done
cat gpurun_out/r2_exp6.jsonl
timeout 300 python scripts/shard_iters.py 8 > gpurun_out/r2_shard_iters_n8.txt 2>&1; tail -20 gpurun_out/r2_shard_iters_n8.txt
