timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_final_cfg4.json 2> gpurun_out/r2_bench_final_cfg4.err; echo "cfg4 rc=$?"; tail -n 2 gpurun_out/r2_bench_final_cfg4.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --workload cfg2 > gpurun_out/r2_bench_final_cfg2.json 2> gpurun_out/r2_bench_final_cfg2.err; echo "cfg2 rc=$?"
timeout 900 python bench.py --workload cfg5 > gpurun_out/r2_bench_final_cfg5.json 2> gpurun_out/r2_bench_final_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
def last(f): return json.loads(open(f).read().strip().splitlines()[-1])
d=last('gpurun_out/r2_bench_final_cfg4.json')
print('cfg4', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_of_burst_peak'), d['step_roofline'], d['e2e']['value'], d['e2e']['resident_bank']['value'], d['parity_check']['ok'], d['cpu_baseline']['value'], d['clocks'])
print(json.dumps(d['extra'])[:1200])
d=last('gpurun_out/r2_bench_final_reference.json'); print('ref', d['value'], d['cpu_baseline']['cores'])
d=last('gpurun_out/r2_bench_final_cfg2.json'); print('cfg2', d['value'], d['ms_per_step'], d['torch_gpu'], d['parity_check']['ok'], d.get('cpu_baseline'))
d=last('gpurun_out/r2_bench_final_cfg5.json'); print('cfg5', d['value'], d['ms_per_step'], d['parity_check']['ok'], {k:(round(v['ms'],3), round(v.get('torch_gpu_ms',0),2)) for k,v in (d.get('sweep') or d.get('extra') or {}).items()} if isinstance((d.get('sweep') or d.get('extra')), dict) else None)
PY
