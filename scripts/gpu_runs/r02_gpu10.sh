#!/bin/bash
# 1-GPU: validate the new recurrence / rank-K kernels, A/B them, in-situ timeline
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r02_pytest_gpu_f.log
grep -E "passed|failed|^E  |FAILED" gpurun_out/r02_pytest_gpu_f.log | head
bash scripts/ab_bench.sh "DRN_QE_FWD2=0" "DRN_QE_FWD2=1" 2>&1 | tee gpurun_out/r02_ab_fwd2.log
timeout 300 python scripts/insitu_timeline.py > gpurun_out/r02_insitu_g.json 2> gpurun_out/r02_insitu_g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_insitu_g.json'))
print(d['step_us_with_stamps'], d['stamp_overhead_us'])
print({k:v for k,v in d['by_call_us'].items() if v>20})
print([x for x in d['sequence_us'] if x[0] in ('qe_forward','qe_backward','gates_bwd','gates')])
PY
