set -x
(timeout 1800 python -m pytest tests/test_loss_gpu.py tests/test_gemm_gpu.py tests/test_retrieval_gpu.py tests/test_simtopk_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short 2>&1 | tail -80) > gpurun_out/r2_gputest3.log
tail -30 gpurun_out/r2_gputest3.log
timeout 900 python tools/loss_accuracy.py > gpurun_out/r2_loss_accuracy.log 2>&1; tail -20 gpurun_out/r2_loss_accuracy.log
MCLST_LOSS_LEAN=0 timeout 900 python tools/loss_accuracy.py > gpurun_out/r2_loss_accuracy_general.log 2>&1
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 rc=$?"
tail -n 3 gpurun_out/r2_bench_cfg5.err gpurun_out/r2_bench_cfg2.err gpurun_out/r2_bench_cfg4.err
