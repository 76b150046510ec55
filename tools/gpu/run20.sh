for R in 2 3 4; do echo "== ring $R"; MCLST_TF32_RING=$R timeout 300 python tools/gemm_bench.py 2>&1 | tail -9; done > gpurun_out/gemm_bench_v3.log 2>&1
cat gpurun_out/gemm_bench_v3.log
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
