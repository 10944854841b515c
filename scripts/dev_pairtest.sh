mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 scripts/dp_check.py > gpurun_out/r2_dp_check_2gpu.txt 2>&1; echo rc=$?; tail -2 gpurun_out/r2_dp_check_2gpu.txt
