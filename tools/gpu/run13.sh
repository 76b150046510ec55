# state check of HEAD + data for the next decisions: full GPU suite, cfg2 graph timeline, default bench,
# source-level ncu of rerank
set -x
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -30) > gpurun_out/r2_gputest13.log
tail -6 gpurun_out/r2_gputest13.log
timeout 300 python tools/cfg2_graph_timeline.py 2>&1 | tail -18 > gpurun_out/r2_cfg2_graph_timeline.log; cat gpurun_out/r2_cfg2_graph_timeline.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/r2_bench13_cfg4.json 2> gpurun_out/r2_bench13_cfg4.err; echo "cfg4 rc=$?"
tail -n 3 gpurun_out/r2_bench13_cfg4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rerank_kernel' -s 3 -c 1 \
    -o gpurun_out/r2_prof_rerank python bench.py --steps 1 --warmup 3 --no-e2e --no-extra --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2_ncu_rerank.err; echo "ncu rc=$?"
tail -n 3 gpurun_out/r2_ncu_rerank.err
ls -la gpurun_out/
