#!/bin/bash
# ncu sections (speed of light, memory workload, occupancy) of the HBM- / latency-bound kernels of one step (run under gpurun)
set -u
TAG=${1:-v}
mkdir -p gpurun_out
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --clock-control none \
    -k regex:'col_stats|bn_bwd_apply|bn_relu_apply|head_proj|gate_reduce|pack_conv|unpack_conv|split_planes|linear_small|sgemm_multi|lstm_|qe_attn|pos_|fcos_loss' \
    --launch-skip 80 -c 80 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_small_$TAG.txt 2>&1
echo "ncu small rc=$?"
grep -c "Duration" gpurun_out/ncu_small_$TAG.txt
