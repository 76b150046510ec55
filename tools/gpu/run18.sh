for R in 2 3 4; do echo "== ring $R"; MCLST_TF32_RING=$R timeout 300 python tools/gemm_bench.py 2>&1 | tail -9; done > gpurun_out/gemm_bench_rings.log 2>&1
cat gpurun_out/gemm_bench_rings.log
