set -x
(timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2_gputest15_gemm.log
tail -25 gpurun_out/r2_gputest15_gemm.log
(timeout 900 python -m pytest tests/test_model_gpu.py tests/test_optim_gpu.py tests/test_evaluate_gpu.py -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2_gputest15_model.log
tail -8 gpurun_out/r2_gputest15_model.log
DUMP=gpurun_out/tl_tf32.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_tf32.log; head -12 gpurun_out/tl_tf32.log
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench15_cfg2.json 2> gpurun_out/r2_bench15_cfg2.err; echo "cfg2 rc=$?"; tail -n 3 gpurun_out/r2_bench15_cfg2.err
