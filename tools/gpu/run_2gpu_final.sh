(timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q --tb=short 2>&1 | tail -4) > gpurun_out/r2_dist_pytest_2gpu.log; tail -2 gpurun_out/r2_dist_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r2_scale_n2_final.json 2> gpurun_out/r2_scale_n2_final.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_scale_n2_final.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['config']['parallelism'], d['other_grids'], d['e2e']['value'], d['parity_check']['ok'], d['roofline']['kernels_ms_per_step'], d['clocks'])
PY
