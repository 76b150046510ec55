timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_loss_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
timeout 300 python tools/gemm_bench.py 2>&1 | tail -9
DUMP=gpurun_out/tl_v9.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_v9.log; head -8 gpurun_out/tl_v9.log
