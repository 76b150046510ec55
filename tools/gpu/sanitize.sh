# compute-sanitizer over reduced-size GPU tests (memcheck / racecheck / synccheck
# on every kernel).  Writes one log per tool under gpurun_out/; tools/summarize_sanitizer.py turns
# them into profiles/r2_sanitizer.md.
set -x
SEL='tests/test_gemm_gpu.py::test_matmul_transposes tests/test_gemm_gpu.py::test_matmul_nt_epilogue tests/test_gemm_gpu.py::test_matmul_rows_of_very_different_scale_and_outliers tests/test_loss_gpu.py::test_loss_vs_reference_golden tests/test_simtopk_gpu.py tests/test_metrics_gpu.py tests/test_optim_gpu.py::test_dense_kernel_matches_torch_adam tests/test_optim_gpu.py::test_lazy_rows_equal_dense_bit_for_bit tests/test_retrieval_gpu.py::test_pm1_known_answer_against_reference_values tests/test_retrieval_gpu.py::test_staged_without_bound_equals_single_call_and_extreme_bounds tests/test_retrieval_gpu.py::test_q1_squeeze_quirk'
for TOOL in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $TOOL --target-processes all --print-limit 50 --error-exitcode 0 \
      --log-file gpurun_out/r2_sanitizer_$TOOL.log \
      python -m pytest $SEL -m gpu -q -x --tb=line -p no:cacheprovider > gpurun_out/r2_sanitizer_${TOOL}_pytest.log 2>&1
  echo "$TOOL rc=$?"
  tail -n 3 gpurun_out/r2_sanitizer_${TOOL}_pytest.log
  grep -c "=========" gpurun_out/r2_sanitizer_$TOOL.log | head -1
  tail -n 5 gpurun_out/r2_sanitizer_$TOOL.log
done
