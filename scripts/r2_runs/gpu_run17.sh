#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest17.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest17.log
tail -3 gpurun_out/r2_pytest17.log
timeout 300 python scripts/shard_iters.py 1 2>&1 | grep -E "setup|it  0|TOTAL" | tail -3
timeout 300 python scripts/exp.py c4 cone2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4', d['ms_frame_1'], d['ms_split_1'])"
