#!/bin/bash
N=8
export RBRT_DEBUG_NO_GATHER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_c3_n8_nogather.json 2> gpurun_out/r2_bench_c3_n8_nogather.err; echo "bench exit $?"
python - <<PY
import json
b=json.load(open('gpurun_out/r2_bench_c3_n8_nogather.json'))
print('NO GATHER value',round(b['value']),'ms/step',round(b['ms_per_step'],3),'fpb1',round(b['frames_per_batch_1']['ms_per_step'],3),'single',round(b['single_frame']['ms_per_step'],3),'e2e',round(b['e2e']['value']),round(b['e2e']['ms_per_step'],3))
PY
tail -3 gpurun_out/r2_bench_c3_n8_nogather.err | cut -c1-300
