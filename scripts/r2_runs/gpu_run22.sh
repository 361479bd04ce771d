#!/bin/bash
SPP=8 timeout 300 python scripts/c4_diff.py 2>&1 | tail -16
echo "--- no primary cull"; RBRT_NO_PRIMARY_CULL=1 SPP=8 timeout 300 python scripts/c4_diff.py 2>&1 | tail -3
