timeout 1200 python -m pytest tests/test_retrieval_gpu.py tests/test_simtopk_gpu.py tests/test_evaluate_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_bench25_cfg4.json 2> gpurun_out/r2_bench25_cfg4.err; echo "cfg4 rc=$?"; tail -n 3 gpurun_out/r2_bench25_cfg4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench25_cfg4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['kernels_ms_per_step'], d['roofline']['frac'], d['step_roofline'], d['e2e']['value'], d['e2e']['resident_bank']['value'], d['parity_check']['ok'], d['path_counters'], d['clocks'])
PY
