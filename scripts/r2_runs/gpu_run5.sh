#!/bin/bash
# N-GPU run: one-process check (NCCL + PEER transports, CLI --gpus), then bench.py under torchrun at N with 4 and with 2 frames per batch
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus_n$N.txt
timeout 600 python scripts/multi_gpu_check.py $N > gpurun_out/r2_multi_check_n$N.json 2> gpurun_out/r2_multi_check_n$N.err; echo "check exit $?"
cat gpurun_out/r2_multi_check_n$N.json; grep "\[multi\]" gpurun_out/r2_multi_check_n$N.err | head -40
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "bench exit $?"
cat gpurun_out/r2_bench_c3_n$N.json; grep "e2e rank 0" gpurun_out/r2_bench_c3_n$N.err | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --frames-per-batch 2 --no-cpu > gpurun_out/r2_bench_c3_n${N}_fpb2.json 2> gpurun_out/r2_bench_c3_n${N}_fpb2.err; echo "bench fpb2 exit $?"
cat gpurun_out/r2_bench_c3_n${N}_fpb2.json
