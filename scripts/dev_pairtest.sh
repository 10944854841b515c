mkdir -p gpurun_out
timeout 120 python scripts/ppo_graph_probe.py > gpurun_out/ppo_graph.txt 2>&1; tail -4 gpurun_out/ppo_graph.txt
