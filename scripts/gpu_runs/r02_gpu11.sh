#!/bin/bash
# 1-GPU: ncu of the query-encoder kernels (new recurrence), stage-2 test, multi-seed R@1 parity
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multistep_gpu.py tests/test_query_gpu.py -q 2>&1 | tail -4
timeout 600 ncu --section SpeedOfLight --section Occupancy --section LaunchStats --section WarpStateStats --clock-control none \
    -k regex:'lstm_|sgemm_multi|linear_small|qe_' --launch-skip 40 -c 40 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/r02_ncu_qe.txt 2>&1
grep -E "^  [a-z_:<>0-9A-Za-z, ]+\(|Duration|Registers Per|Achieved Occupancy|Grid Size|Compute \(SM\) Throughput|Memory Throughput" gpurun_out/r02_ncu_qe.txt | head -150
timeout 1500 python scripts/r1_parity_seeds.py --seeds 6 --steps 35 --eval-batches 16 > gpurun_out/r02_r1_parity_seeds.json 2> gpurun_out/r02_r1_parity_seeds.err
tail -8 gpurun_out/r02_r1_parity_seeds.err; python -c "
import json; d=json.load(open('gpurun_out/r02_r1_parity_seeds.json')); print(json.dumps(d['R@1'])); print(d['verdict'])"
