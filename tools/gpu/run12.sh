set -x
(timeout 1800 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_loss_gpu.py tests/test_optim_gpu.py -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2_gputest12.log
tail -6 gpurun_out/r2_gputest12.log
timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -16 > gpurun_out/r2_cfg2_graph_timeline.log; cat gpurun_out/r2_cfg2_graph_timeline.log
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"
tail -n 2 gpurun_out/r2_bench_cfg2.err gpurun_out/r2_bench_cfg5.err
