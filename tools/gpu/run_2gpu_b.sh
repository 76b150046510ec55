set -x
N=2
(timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2_dist_pytest_${N}gpu.log
tail -5 gpurun_out/r2_dist_pytest_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/r2_scale_n${N}_b0.json 2> gpurun_out/r2_scale_n${N}_b0.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-extra --no-numa > gpurun_out/r2_scale_n${N}_b0_nonuma.json 2> gpurun_out/r2_scale_n${N}_b0_nonuma.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r2_scale_n${N}_cfg5.json 2> gpurun_out/r2_scale_n${N}_cfg5.err; echo "cfg5 rc=$?"
tail -n 2 gpurun_out/r2_scale_n${N}_*.err
