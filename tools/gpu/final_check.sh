# what the driver runs at round end, on one B200: GPU test suite, smoke(), default bench
set -x
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -4) > gpurun_out/final_gputest.log; tail -2 gpurun_out/final_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -n 2 gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['roofline']['frac'], d['roofline'].get('traffic'), d['e2e']['value'], d['e2e']['resident_bank']['value'], d['parity_check']['ok'], d['cpu_baseline']['value'], d['clocks'])
x=d['extra']; print({k:round(v['ms'],3) for k,v in x['contrastive_loss_soft_fwd_bwd'].items()}, x['train_step_cfg2_B1024']['G1000'].get('graph_ms_per_step'))
PY
