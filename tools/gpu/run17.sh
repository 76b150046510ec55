set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 6 -c 1 -o gpurun_out/r2_prof_tf32 python tools/gemm_bench.py > /dev/null 2> gpurun_out/ncu_tf32.err; echo rc=$?
tail -3 gpurun_out/ncu_tf32.err
