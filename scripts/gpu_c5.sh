#!/bin/bash
# C5 (5 242 880 triangles, 3840x2160, 1024 spp) on N GPUs, sample-range sharded, ncclReduce of the accumulators inside the library
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err; echo "bench exit $?"
python - <<PY
import json
b=json.load(open('gpurun_out/r2_bench_c5_n$N.json'))
print('C5 N=$N value',round(b['value']),'ms/step',round(b['ms_per_step'],2),'samples/s',b['samples_per_s'],'single',round(b['single_frame']['ms_per_step'],2),'e2e',round(b['e2e']['value']),round(b['e2e']['ms_per_step'],2),b.get('image_check'))
PY
tail -3 gpurun_out/r2_bench_c5_n$N.err | cut -c1-300
