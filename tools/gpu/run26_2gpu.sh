for QB in 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$QB bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --no-e2e --queries 16384 --query-blocks $QB > gpurun_out/r2_tune_n2_q16k_qb$QB.json 2> gpurun_out/r2_tune_n2_q16k_qb$QB.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_tune_n2_q16k_qb$QB.json').read().strip().splitlines()[-1])
print('QB$QB', d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['parity_check']['ok'])
PY
done
timeout 600 python -m pytest tests/test_retrieval_gpu.py -m gpu -q --tb=short -x -k "blocked" 2>&1 | tail -3
