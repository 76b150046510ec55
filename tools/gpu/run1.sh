set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -60) > gpurun_out/r2_gputest1.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err; echo "cfg4 rc=$?"
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "cfg2 rc=$?"
tail -5 gpurun_out/r2_gputest1.log
