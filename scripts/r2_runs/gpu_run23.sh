#!/bin/bash
SPP=1 timeout 600 python scripts/c4_diff.py 2>&1 | tail -24
