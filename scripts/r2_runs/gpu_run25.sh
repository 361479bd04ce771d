#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest25.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest25.log
tail -4 gpurun_out/r2_pytest25.log
timeout 300 python scripts/exp.py c4 final 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4', d['ms_frame_1'], d['ms_split_1'], d['checksum'])"
