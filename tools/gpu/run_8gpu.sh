set -x
nvidia-smi --query-gpu=index,name --format=csv | head -3
N=8
(timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2_dist_pytest_${N}gpu.log
tail -5 gpurun_out/r2_dist_pytest_${N}gpu.log
for SH in 0 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$SH bench.py --gpus $N --steps 5 --warmup 3 --bank-shards $SH > gpurun_out/r2_scale_n${N}_b${SH}.json 2> gpurun_out/r2_scale_n${N}_b${SH}.err; echo "bench N=$N shards=$SH rc=$?"
  tail -n 2 gpurun_out/r2_scale_n${N}_b${SH}.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r2_scale_n${N}_cfg5.json 2> gpurun_out/r2_scale_n${N}_cfg5.err; echo "cfg5 N=$N rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 5 --warmup 3 --no-extra > gpurun_out/r2_scale_n4_b0.json 2> gpurun_out/r2_scale_n4_b0.err; echo "bench N=4 rc=$?"
tail -n 2 gpurun_out/r2_scale_n${N}_cfg5.err gpurun_out/r2_scale_n4_b0.err
