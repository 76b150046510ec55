set -x
(timeout 1800 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60) > gpurun_out/r2_gputest7.log
tail -15 gpurun_out/r2_gputest7.log
(MCLST_LOSS_MN=0 timeout 900 python -m pytest tests/test_loss_gpu.py -m gpu -q --tb=short -k "at_size or shapes or golden" 2>&1 | tail -8) > gpurun_out/r2_gputest7_nomn.log; tail -3 gpurun_out/r2_gputest7_nomn.log
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 rc=$?"
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 rc=$?"
tail -n 3 gpurun_out/r2_bench_cfg2.err gpurun_out/r2_bench_cfg4.err gpurun_out/r2_bench_cfg5.err
bash tools/gpu/sanitize.sh
