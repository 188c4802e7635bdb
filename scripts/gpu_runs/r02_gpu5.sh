#!/bin/bash
set -u
mkdir -p gpurun_out
PB=4 PT=64 timeout 600 python scripts/bwd_debug.py > gpurun_out/r02_bwd_debug_b4_t64.txt 2>&1
grep -E "level|worst|\(b=|dy tower" gpurun_out/r02_bwd_debug_b4_t64.txt
