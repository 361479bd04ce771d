#!/bin/bash
python scripts/ncu_build.py c3 > gpurun_out/r2_ncu_build_plain.log 2>&1 &&
ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_ncu_build_c3.csv python scripts/ncu_build.py c3 > gpurun_out/r2_ncu_build_run.log 2>&1
echo "exit $?"; tail -2 gpurun_out/r2_ncu_build_plain.log
