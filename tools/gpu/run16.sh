set -x
timeout 300 python tools/gemm_bench.py 2>&1 | tail -12 > gpurun_out/gemm_bench.log; cat gpurun_out/gemm_bench.log
DUMP=gpurun_out/tl_tf32.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_tf32.log; head -6 gpurun_out/tl_tf32.log
