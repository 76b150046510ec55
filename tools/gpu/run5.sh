set -x
(timeout 1800 python -m pytest tests/test_loss_gpu.py tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_optim_gpu.py tests/test_retrieval_gpu.py -m gpu -q --tb=short 2>&1 | tail -60) > gpurun_out/r2_gputest5.log
tail -15 gpurun_out/r2_gputest5.log
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 rc=$?"
STEPS=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cfg2_launches.csv python tools/cfg2_one_step.py > /dev/null 2>&1; echo "ncu cfg2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_loss_launches.csv python tools/loss_profile.py 32768 > /dev/null 2>&1; echo "ncu loss rc=$?"
tail -n 3 gpurun_out/r2_bench_cfg5.err gpurun_out/r2_bench_cfg2.err
