timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6
DUMP=gpurun_out/tl_v5.txt timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -17 > gpurun_out/tl_v5.log; head -4 gpurun_out/tl_v5.log
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench27_cfg2.json 2> gpurun_out/r2_bench27_cfg2.err; echo "cfg2 rc=$?"
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench27_cfg5.json 2> gpurun_out/r2_bench27_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench27_cfg2.json').read().strip().splitlines()[-1])
print('cfg2', d['value'], d['ms_per_step'], d.get('torch_gpu',{}).get('graph_ms_per_step'), d['parity_check']['ok'], d['gpu_launches'])
d=json.loads(open('gpurun_out/r2_bench27_cfg5.json').read().strip().splitlines()[-1])
print('cfg5', d['value'], d['ms_per_step'], d['parity_check'])
print(json.dumps(d.get('sweep') or d.get('extra'))[:1500])
PY
