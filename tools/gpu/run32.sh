timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sim_topk_lanes|rerank|weighted_average' -s 9 -c 4 \
    -o gpurun_out/r2b_prof_retrieval python bench.py --steps 1 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2b_ncu_retrieval.err; echo "retrieval rc=$?"
tail -n 2 gpurun_out/r2b_ncu_retrieval.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity > /dev/null 2>&1; echo "launches rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2b_launches_cfg4.csv
