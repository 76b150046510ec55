for F in 0 1; do for R in 2 4; do echo "== fence $F ring $R"; MCLST_TF32_FENCE=$F MCLST_TF32_RING=$R timeout 300 python tools/gemm_bench.py 2>&1 | tail -9; done; done > gpurun_out/gemm_bench_fence.log 2>&1
cat gpurun_out/gemm_bench_fence.log
MCLST_TF32_FENCE=1 MCLST_TF32_RING=4 timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
