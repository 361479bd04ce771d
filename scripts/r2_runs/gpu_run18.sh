#!/bin/bash
N=${1:-8}
timeout 600 python scripts/multi_gpu_check.py $N > gpurun_out/r2_multi_check_n${N}_threads.json 2> gpurun_out/r2_multi_check_n${N}_threads.err; echo "check exit $?"
RBRT_NO_ENQUEUE_THREADS=1 timeout 600 python scripts/multi_gpu_check.py $N > gpurun_out/r2_multi_check_n${N}_nothreads.json 2> gpurun_out/r2_multi_check_n${N}_nothreads.err; echo "check exit $?"
python - <<PY
import json
for t in ('threads','nothreads'):
    d=json.load(open('gpurun_out/r2_multi_check_n${N}_%s.json' % t))
    for k in ('nccl','peer'):
        e=d[k]; print(t,k,'ms/frame',round(e['ms_per_frame_host_clock'],3),'create',round(e['scene_create_ms'],2),e['u8_identical_to_one_gpu'],e['hdr_identical_to_one_gpu'],e['sample_shards_max_rel_err'])
    print(t,'cli',d['cli']['png_identical'],d['cli']['stderr_tail'], 'one gpu', d['one_gpu']['ms_per_frame_host_clock'])
PY
grep "\[multi\]" gpurun_out/r2_multi_check_n${N}_threads.err | head -11
