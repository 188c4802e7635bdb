#!/bin/bash
# 1-GPU: host-lengths sync hypothesis A/B, full tests, default bench
set -u
mkdir -p gpurun_out
bash scripts/ab_bench.sh "DRN_BENCH_HOST_LENGTHS=1" "DRN_BENCH_HOST_LENGTHS=0" 2>&1 | tee gpurun_out/r02_ab_hostlen.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_g.log
grep -E "passed|failed|FAILED" gpurun_out/r02_pytest_gpu_g.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_g.json'))
print('value %.0f step %.3f e2e %.0f (%.3f) sustained %.3f frac_sus %.3f frac_burst %.3f roofline %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['sustained']['ms_per_step'],d['sustained']['step_frac_of_sustained_peak_issued'],d['roofline']['step_frac_of_burst_peak_issued'],d['roofline']['frac']))
print(d['cpu_baseline']); print(d['extra'].get('fast_mode')); print([ (x['stage'],x['ms_per_step_incl_optimizer']) for x in d['extra']['configs[2]']]); print([(x['T'],x['pairs_per_s_forward_postprocess_metric_on_device']) for x in d['extra']['configs[4]']])"
