timeout 300 python tools/gemm_bench.py 2>&1 | tail -9 > gpurun_out/gemm_bench_elect.log; cat gpurun_out/gemm_bench_elect.log
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_loss_gpu.py tests/test_simtopk_gpu.py tests/test_retrieval_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/r2_bench31_cfg4.json 2> gpurun_out/r2_bench31_cfg4.err; echo "cfg4 rc=$?"; tail -n 3 gpurun_out/r2_bench31_cfg4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench31_cfg4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['roofline']['frac'], d['parity_check']['ok'], d['clocks'])
PY
DUMP=gpurun_out/tl_v7.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_v7.log; head -7 gpurun_out/tl_v7.log
