#!/bin/bash
for n in 8 2 1; do
  for prio in 0 1; do
    for depth in 2 3; do
      RBRT_BUILD_PRIORITY=$prio DEPTH=$depth timeout 200 python scripts/e2e_shard_time.py c3 $n 12 2>&1 | tail -1
    done
  done
  DEPTHS=2 timeout 200 python scripts/pipe_time.py c3 $n 12 2>&1 | tail -1
done
