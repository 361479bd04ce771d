#!/bin/bash
# N-GPU: bench.py under torchrun (process per GPU, NCCL inside the library), then the one-process check
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_n${N}_b.json 2> gpurun_out/r2_bench_c3_n${N}_b.err; echo "bench exit $?"
python - <<PY
import json
b=json.load(open('gpurun_out/r2_bench_c3_n${N}_b.json'))
print('value',round(b['value']),'ms/step',round(b['ms_per_step'],3),'single',round(b['single_frame']['ms_per_step'],3),'one lane',round(b['single_frame']['ms_per_step_one_lane'],3),'fpb1',round(b['frames_per_batch_1']['ms_per_step'],3),'e2e',round(b['e2e']['value']),round(b['e2e']['ms_per_step'],3),b['image_check'])
PY
grep "e2e rank 0" gpurun_out/r2_bench_c3_n${N}_b.err | tail -1
timeout 600 python scripts/multi_gpu_check.py $N > gpurun_out/r2_multi_check_n${N}_b.json 2> gpurun_out/r2_multi_check_n${N}_b.err; echo "check exit $?"
cat gpurun_out/r2_multi_check_n${N}_b.json; grep "\[multi\]" gpurun_out/r2_multi_check_n${N}_b.err | head -12
