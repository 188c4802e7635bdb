#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/grad_error_table.py --golden > gpurun_out/r02_grad_errors_golden.json 2> gpurun_out/r02_grad_errors_golden.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_grad_errors_golden.json'))
for k,v in d['stages'].items():
    print(k, 'max_err %.2e max err/sens %.2f max(err-4sens) %.2e'%(v['max_err'], v['max_err_over_sens'], v['max_err_minus_4sens']))
    rows=sorted([(r['err']-4*r['sens'],kk,r['err'],r['sens']) for kk,r in v['per_tensor'].items() if 'err' in r], reverse=True)[:4]
    for x in rows: print('     %-46s err %.2e sens %.2e'%(x[1],x[2],x[3]))
PY
timeout 900 python -m pytest tests/test_multistep_gpu.py tests/test_metric_gpu.py -q -s 2>&1 | grep -E "five steps|passed|failed|^E  |Error" | head -20
