timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_optim_gpu.py tests/test_evaluate_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -12
timeout 300 python tools/gemm_bench.py 2>&1 | tail -9 > gpurun_out/gemm_bench_mixed.log; cat gpurun_out/gemm_bench_mixed.log
DUMP=gpurun_out/tl_mixed.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_mixed.log; head -12 gpurun_out/tl_mixed.log
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench24_cfg2.json 2> gpurun_out/r2_bench24_cfg2.err; echo "cfg2 rc=$?"; tail -n 2 gpurun_out/r2_bench24_cfg2.err
