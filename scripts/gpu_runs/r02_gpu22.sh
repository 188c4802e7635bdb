#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/gemm_trace.py > gpurun_out/r02_gemm_trace.json 2> gpurun_out/r02_gemm_trace.err
tail -3 gpurun_out/r02_gemm_trace.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_gemm_trace.json'))
keys=["ctas","tiles","balanced_schedule","last_cta_entered","prologue_done_median","first_operands_landed_median","last_mma_issued_median","last_mma_issued_max","last_accumulator_complete_max","drained_median","drained_max","exit_max","since_previous_traced_exit"]
print("%-26s"%"launch"+" ".join("%8s"%k[:8] for k in keys))
for r in d["launches"]:
    print("%-26s"%r["launch"]+" ".join("%8s"%r[k] for k in keys))
PY
