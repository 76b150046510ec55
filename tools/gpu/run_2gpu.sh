set -x
nvidia-smi --query-gpu=index,name --format=csv
N=${N:-2}
(timeout 1500 python -m pytest tests/test_distributed_gpu.py -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2_dist_pytest_${N}gpu.log
tail -20 gpurun_out/r2_dist_pytest_${N}gpu.log
for SH in 0 $N 1; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-extra --bank-shards $SH > gpurun_out/r2_scale_n${N}_b${SH}.json 2> gpurun_out/r2_scale_n${N}_b${SH}.err; echo "bench N=$N shards=$SH rc=$?"
  tail -n 2 gpurun_out/r2_scale_n${N}_b${SH}.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r2_scale_n${N}_cfg5.json 2> gpurun_out/r2_scale_n${N}_cfg5.err; echo "cfg5 N=$N rc=$?"
tail -n 2 gpurun_out/r2_scale_n${N}_cfg5.err
