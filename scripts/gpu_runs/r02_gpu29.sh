#!/bin/bash
# r02: full GEMM + model tests after the epilogue / schedule changes, then compute-sanitizer memcheck on the small cases
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_query_gpu.py tests/test_multistep_gpu.py -q -x 2>&1 | tail -3
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x \
    "tests/test_gemm_gpu.py::test_fused_batchnorm_partial_sums" \
    "tests/test_gemm_gpu.py::test_epilogue_options" \
    "tests/test_gemm_gpu.py::test_group_mixed_forms_and_split_slices" \
    "tests/test_gemm_gpu.py::test_hybrid_schedule_splits_only_the_last_wave" \
    "tests/test_query_gpu.py::test_sgemm_forms" \
    "tests/test_model_gpu.py::test_gradients_match_oracle[s1_train_b4_t32]" \
    "tests/test_model_gpu.py::test_gradients_match_oracle[s1_train_b4_t32_charades]" \
    "tests/test_model_gpu.py::test_gradients_match_oracle[s2_train_b4_t32_crafted]" \
    "tests/test_model_gpu.py::test_forward_matches_oracle_and_reference_golden[s3_eval_b3_t64_crafted]" \
    "tests/test_model_gpu.py::test_token_width_buckets" > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" gpurun_out/r02_sanitizer_memcheck.log | head -10
bash scripts/ab_bench.sh "DRN_SCHEDULE=hybrid" "DRN_SCHEDULE=static" 2>&1 | tee gpurun_out/r02_ab_hybrid2.log
