timeout 300 python tools/gemm_bench.py 2>&1 | tail -9 > gpurun_out/gemm_bench_v4.log; cat gpurun_out/gemm_bench_v4.log
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
DUMP=gpurun_out/tl_tf32.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_tf32.log; head -8 gpurun_out/tl_tf32.log
