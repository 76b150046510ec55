set -x
nvidia-smi --query-gpu=index,name --format=csv | head -3
N=8
(timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2_dist_pytest_${N}gpu.log
tail -5 gpurun_out/r2_dist_pytest_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_n${N}_final.json 2> gpurun_out/r2_scale_n${N}_final.err; echo "bench N=$N rc=$?"
tail -n 2 gpurun_out/r2_scale_n${N}_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --steps 5 --warmup 3 --no-extra > gpurun_out/r2_scale_n4_final.json 2> gpurun_out/r2_scale_n4_final.err; echo "bench N=4 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r2_scale_n${N}_cfg5_final.json 2> gpurun_out/r2_scale_n${N}_cfg5_final.err; echo "cfg5 N=$N rc=$?"
tail -n 2 gpurun_out/r2_scale_n4_final.err gpurun_out/r2_scale_n${N}_cfg5_final.err
python - <<'PY'
import json
for f in ['gpurun_out/r2_scale_n8_final.json','gpurun_out/r2_scale_n4_final.json','gpurun_out/r2_scale_n8_cfg5_final.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['value'], d['config'].get('parallelism'), d.get('other_grids'), (d.get('e2e') or {}).get('value'), (d.get('parity_check') or {}).get('ok'))
    except Exception as e:
        print(f, 'ERR', e)
PY
