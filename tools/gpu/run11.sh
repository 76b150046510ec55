set -x
(timeout 1800 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -30) > gpurun_out/r2_gputest11.log
tail -6 gpurun_out/r2_gputest11.log
timeout 300 python tools/cfg2_graph_timeline.py > gpurun_out/r2_cfg2_graph_timeline.log 2>&1; cat gpurun_out/r2_cfg2_graph_timeline.log | tail -20
bash tools/gpu/ncu_full.sh
