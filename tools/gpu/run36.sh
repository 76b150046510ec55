MCLST_LIB_NAME=libmclst_dbg.so timeout 120 python tools/gemm_timing.py 1024 1000 1000 2>&1 | tail -9
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_loss_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -2
timeout 300 python tools/gemm_bench.py 2>&1 | tail -9 | head -4
