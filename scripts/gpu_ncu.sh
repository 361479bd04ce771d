#!/bin/bash
# ncu evidence for the roofline record: (1) a metrics pass over every launch of one warm C3 frame for tile-shard denominators 1, 2, 4, 8;
# (2) one --set full capture (source view) of the first trace / shade / generate launches of the full frame.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
for n in 1 2 4 8; do
  python scripts/ncu_frame.py c3 $n > gpurun_out/r2_ncu_plain_n$n.log 2>&1 &&
  ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r2_ncu_metrics_c3_n$n.csv python scripts/ncu_frame.py c3 $n > gpurun_out/r2_ncu_run_n$n.log 2>&1
  echo "ncu metrics n=$n exit $?"; tail -1 gpurun_out/r2_ncu_plain_n$n.log
done
python scripts/ncu_frame.py c3 1 > gpurun_out/r2_ncu_plain_full.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_trace|k_shade|k_generate" -c 5 -o gpurun_out/r2_prof_c3 python scripts/ncu_frame.py c3 1 > gpurun_out/r2_ncu_run_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
