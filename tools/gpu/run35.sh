MCLST_LIB_NAME=libmclst_dbg.so timeout 120 python tools/gemm_timing.py 1024 1000 1000 2>&1 | tail -9
MCLST_LIB_NAME=libmclst_dbg.so timeout 120 python tools/gemm_timing.py 1024 256 1024 2>&1 | tail -9
