#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -4
for cfg in "DRN_EPI_T=1" "DRN_EPI_T=0"; do
echo "---- [$cfg]"
env $cfg timeout 300 python scripts/gemm_trace.py > gpurun_out/r02_gemm_trace_$cfg.json 2> gpurun_out/r02_gemm_trace.err
tail -3 gpurun_out/r02_gemm_trace.err
python - gpurun_out/r02_gemm_trace_$cfg.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
keys=["ctas","tiles","first_operands_landed_median","last_mma_issued_median","last_mma_issued_max","last_accumulator_complete_max","drained_median","drained_max","exit_max"]
print("%-26s"%"launch"+" ".join("%8s"%k[:8] for k in keys))
for r in d["launches"]:
    print("%-26s"%r["launch"]+" ".join("%8s"%r[k] for k in keys))
print("sum of exit_max", sum(r["exit_max"] for r in d["launches"]))
PY
done
bash scripts/ab_bench.sh "DRN_EPI_T=1" "DRN_EPI_T=0" 2>&1 | tee gpurun_out/r02_ab_epit.log
