timeout 300 python tools/gemm_bench.py 2>&1 | tail -9 | head -4
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -2
DUMP=gpurun_out/tl_v8.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_v8.log; head -6 gpurun_out/tl_v8.log
