#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -3
bash scripts/ab_bench.sh "" 2>&1 | tee gpurun_out/r02_ab_epi8.log
timeout 300 python scripts/insitu_timeline.py > gpurun_out/r02_insitu_epi8.json 2> gpurun_out/r02_insitu_epi8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_insitu_epi8.json'))
print([ (n, round(v-6.9,1)) for n,v in d['sequence_us'] if n.startswith('drn_gemm')])
PY
