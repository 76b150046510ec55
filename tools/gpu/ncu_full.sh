# ncu --set full captures of the round-2 top kernels (one GPU; each kernel is replayed ~40 times)
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sim_topk_lanes|rerank|weighted_average' -s 9 -c 4 \
    -o gpurun_out/r2_prof_retrieval python bench.py --steps 1 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2_ncu_retrieval.err; echo "retrieval rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tn|grad_tiles|soft_pass2|col_lse' -s 14 -c 7 \
    -o gpurun_out/r2_prof_loss python tools/loss_profile.py 32768 > /dev/null 2> gpurun_out/r2_ncu_loss.err; echo "loss rc=$?"
ls -la gpurun_out/*.ncu-rep
tail -n 3 gpurun_out/r2_ncu_retrieval.err gpurun_out/r2_ncu_loss.err
