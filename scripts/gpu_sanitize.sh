#!/bin/bash
# compute-sanitizer memcheck over the kernels changed in r01 v11..v16 (small cases; run under gpurun)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py::test_fused_batchnorm_partial_sums tests/test_model_gpu.py::test_postprocess_kernel_matches_oracle -q 2>&1 | tail -3
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x \
    "tests/test_gemm_gpu.py::test_fused_batchnorm_partial_sums" \
    "tests/test_query_gpu.py::test_sgemm_forms" \
    "tests/test_model_gpu.py::test_gradients_match_oracle[s1_train_b4_t32]" \
    "tests/test_model_gpu.py::test_forward_matches_oracle_and_reference_golden[s3_eval_b3_t64_crafted]" \
    "tests/test_model_gpu.py::test_token_width_buckets" > gpurun_out/sanitizer_memcheck_v16.log 2>&1
echo "sanitizer rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/sanitizer_memcheck_v16.log | head -10
