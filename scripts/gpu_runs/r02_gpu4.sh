#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_d.log
grep -E "five steps|passed|failed|^E  |Error" gpurun_out/r02_pytest_gpu_d.log | head -30
timeout 600 python scripts/adam_divergence_probe.py > gpurun_out/r02_adam_probe.json 2> gpurun_out/r02_adam_probe.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_adam_probe.json'))
print('rms du all: cuda %.3e ctl %.3e'%(d['rms_du_all_cuda'],d['rms_du_all_ctl']))
rows=sorted(d['per_tensor'].items(), key=lambda kv:-kv[1]['rms_du_cuda']*kv[1]['numel']**0.5)
for k,r in rows[:25]:
    print('%-48s n=%9d rel c/p %.1e %.1e  du c/p %.2e %.2e  flips c/p %.3f %.3f  zero ref/cuda %.3f %.3f <1e-8 %.3f'%(k,r['numel'],r['rel_cuda'],r['rel_ctl'],r['rms_du_cuda'],r['rms_du_ctl'],r['sign_flips_cuda'],r['sign_flips_ctl'],r['frac_exact_zero_ref'],r['frac_exact_zero_cuda'],r['frac_abs_lt_1e-8_ref']))
PY
