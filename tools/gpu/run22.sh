for D in 0 1 2; do echo "== dbg $D"; MCLST_TF32_DBG=$D timeout 300 python tools/gemm_bench.py 2>&1 | tail -9 | head -5; done > gpurun_out/gemm_bench_dbg.log 2>&1
cat gpurun_out/gemm_bench_dbg.log
