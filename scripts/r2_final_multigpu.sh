#!/bin/bash
# Round-2 multi-GPU evidence: N = $1 GPUs of one node (gpurun --gpus N).
N=${1:-2}
mkdir -p gpurun_out
P=$((29500 + N))
if [ "$N" = "2" ]; then
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+50)) scripts/dp_check.py > gpurun_out/r2_dp_check_2gpu.txt 2>&1; tail -2 gpurun_out/r2_dp_check_2gpu.txt
fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; head -c 300 gpurun_out/r2_bench_${N}gpu.json; echo
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+100)) bench.py --config 5 --gpus $N > gpurun_out/r2_config5_${N}gpu.json 2> gpurun_out/r2_config5_${N}gpu.err; cat gpurun_out/r2_config5_${N}gpu.json
