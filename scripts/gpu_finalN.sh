#!/bin/bash
# the driver's bench command at N GPUs (process per GPU), both arms
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "bench exit $?"
python - <<PY
import json
b=json.load(open('gpurun_out/r2_bench_c3_n$N.json'))
print('N=$N value',round(b['value']),'ms/step',round(b['ms_per_step'],3),'single',round(b['single_frame']['ms_per_step'],3),'one lane',round(b['single_frame']['ms_per_step_one_lane'],3),'fpb1',round(b['frames_per_batch_1']['ms_per_step'],3),'e2e',round(b['e2e']['value']),round(b['e2e']['ms_per_step'],3),'bcast',b['e2e']['ms_per_step_scene_broadcast'],b['image_check']['bit_identical'],'dram_frac',b['roofline']['dram_frac'],'frac',b['roofline']['frac'], b['clocks'])
PY
grep "e2e rank 0" gpurun_out/r2_bench_c3_n$N.err | tail -2
