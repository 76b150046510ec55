set -x
(timeout 900 python -m pytest tests/test_model_gpu.py tests/test_optim_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -30) > gpurun_out/r2_gputest14.log
tail -8 gpurun_out/r2_gputest14.log
DEFER=0 MCLST_PARALLEL_BRANCHES=0 DUMP=gpurun_out/tl_base.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_base.log; head -3 gpurun_out/tl_base.log
DUMP=gpurun_out/tl_new.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_new.log; head -3 gpurun_out/tl_new.log
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench14_cfg2.json 2> gpurun_out/r2_bench14_cfg2.err; echo "cfg2 rc=$?"; tail -n 3 gpurun_out/r2_bench14_cfg2.err
