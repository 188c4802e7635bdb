#!/bin/bash
# LPT schedule + finer conv1/lateral K-splits + LSTM barrier (red.release, Hprev fused): tests, A/B, timeline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_query_gpu.py tests/test_model_gpu.py tests/test_multistep_gpu.py -q -x 2>&1 | tail -4
bash scripts/ab_bench.sh "" "DRN_LPT=0" "DRN_SIDE=1" 2>&1 | tee gpurun_out/r02_ab_lpt.log
timeout 300 python scripts/insitu_timeline.py > gpurun_out/r02_insitu_lpt.json 2> gpurun_out/r02_insitu_lpt.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_insitu_lpt.json'))
print([(n, v) for n,v in d['sequence_us'] if n.startswith('drn_gemm') or n.startswith('qe_')])
PY
