#!/bin/bash
for n in 8; do
  for ho in 1 0; do for dep in 1 0; do
    HOSTOUT=$ho DEP=$dep DEPTH=2 timeout 200 python scripts/e2e_shard_time.py c3 $n 16 2>&1 | tail -1
  done; done
  HOSTOUT=0 DEP=1 DEPTH=3 timeout 200 python scripts/e2e_shard_time.py c3 $n 16 2>&1 | tail -1
  HOSTOUT=0 DEP=1 DEPTH=4 timeout 200 python scripts/e2e_shard_time.py c3 $n 16 2>&1 | tail -1
done
