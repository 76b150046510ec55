MCLST_LIB_NAME=libmclst_dbg.so timeout 120 python tools/tf32_timing.py 1024 256 256 > gpurun_out/tf32_timing_a.log 2>&1; cat gpurun_out/tf32_timing_a.log
MCLST_LIB_NAME=libmclst_dbg.so timeout 120 python tools/tf32_timing.py 1024 512 1024 2 > gpurun_out/tf32_timing_b.log 2>&1; head -24 gpurun_out/tf32_timing_b.log
