#!/bin/bash
# r02 second GPU pass: whole GPU suite, both bench arms with the new keys, gradient error table
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_b.log
tail -12 gpurun_out/r02_pytest_gpu_b.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_b.json 2> gpurun_out/r02_bench_ref_b.err
tail -c 1500 gpurun_out/r02_bench_ref_b.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/r02_bench_b.json; tail -5 gpurun_out/r02_bench_b.err
timeout 900 python scripts/grad_error_table.py > gpurun_out/r02_grad_errors.json 2> gpurun_out/r02_grad_errors.err
echo "grad table rc=$?"; tail -3 gpurun_out/r02_grad_errors.err
