# compute-sanitizer over the kernels and stream patterns added late in round 2: the 3xTF32 GEMM with
# the in-kernel split, mclst_retrieve (+ resident packed bank), the side-stream loss products, the
# deferred weight gradients / two-branch forward, the re-rank changes.
set -x
SEL='tests/test_gemm_gpu.py::test_matmul_in_kernel_split_route tests/test_model_gpu.py::test_deferred_weight_grads_equal_autograd_route tests/test_retrieval_gpu.py::test_resident_bank_and_pinned_io_equal_one_shot tests/test_retrieval_gpu.py::test_pm1_known_answer_against_reference_values tests/test_loss_gpu.py::test_loss_vs_reference_golden'
for TOOL in memcheck synccheck racecheck; do
  timeout 300 compute-sanitizer --tool $TOOL --target-processes all --print-limit 50 --error-exitcode 0 \
      --log-file gpurun_out/r2b_sanitizer_$TOOL.log \
      python -m pytest $SEL -m gpu -q -x --tb=line -p no:cacheprovider > gpurun_out/r2b_sanitizer_${TOOL}_pytest.log 2>&1
  echo "$TOOL rc=$?"
  tail -n 2 gpurun_out/r2b_sanitizer_${TOOL}_pytest.log
  tail -n 3 gpurun_out/r2b_sanitizer_$TOOL.log
done
