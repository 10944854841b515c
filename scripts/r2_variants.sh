#!/bin/bash
# Parity of the NON-default kernel selections (each is a supported switch, README): full VAE / tensor-core GPU suites.
mkdir -p gpurun_out
run() { echo "== $1"; env $1 timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_vae_gpu.py tests/test_vae_large_gpu.py -q 2>&1 | tail -2; }
{
run "CPB_TC_PAIR=0"
run "CPB_TC3_WGRAD=1"
run "CPB_TC2_WGRAD=1"
run "CPB_TC2=0"
run "CPB_TC_CLUSTER=1"
} > gpurun_out/r2_pytest_gpu_variants.txt 2>&1
cat gpurun_out/r2_pytest_gpu_variants.txt
