#!/bin/bash
# gpu suite + smoke + the N=1 bench lines of the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_final.log
tail -3 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_bench1.sh
