#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_query_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -3
bash scripts/ab_bench.sh "" 2>&1 | tee gpurun_out/r02_ab_ffma2.log
timeout 300 python scripts/insitu_timeline.py > gpurun_out/r02_insitu_ffma2.json 2> gpurun_out/r02_insitu_ffma2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_insitu_ffma2.json'))
print([(n, v) for n,v in d['sequence_us'] if n.startswith('qe_') or n.startswith('gates')])
PY
