#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_multistep_gpu.py 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_c.log
tail -5 gpurun_out/r02_pytest_gpu_c.log
timeout 900 python -m pytest tests/test_multistep_gpu.py -q -s 2>&1 | tail -60 > gpurun_out/r02_pytest_multistep.log
grep -E "five steps|passed|failed|assert|Error" gpurun_out/r02_pytest_multistep.log | head -20
timeout 600 python scripts/configs_bench.py --skip-train --steps 10 --warmup 3 > gpurun_out/r02_configs4.jsonl 2> gpurun_out/r02_configs4.err
cat gpurun_out/r02_configs4.jsonl | cut -c1-900
