#!/bin/bash
# N-GPU run: one-process check (NCCL + PEER transports, CLI --gpus), then bench.py under torchrun at N (process per GPU, NCCL)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus_n$N.txt
timeout 600 python scripts/multi_gpu_check.py $N > gpurun_out/r2_multi_check_n$N.json 2> gpurun_out/r2_multi_check_n$N.err; echo "check exit $?"
cat gpurun_out/r2_multi_check_n$N.json; tail -3 gpurun_out/r2_multi_check_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "bench exit $?"
cat gpurun_out/r2_bench_c3_n$N.json; tail -5 gpurun_out/r2_bench_c3_n$N.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_n1_a.json 2> gpurun_out/r2_bench_c3_n1_a.err; echo "bench1 exit $?"
cat gpurun_out/r2_bench_c3_n1_a.json; tail -3 gpurun_out/r2_bench_c3_n1_a.err
