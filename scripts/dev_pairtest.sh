mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_tc_gpu.py -q -x -k wgrad > gpurun_out/p12_tc.txt 2>&1; tail -2 gpurun_out/p12_tc.txt
if ! grep -q "12 passed" gpurun_out/p12_tc.txt; then exit 1; fi
export CPB_TC3_WGRAD=1
timeout 120 python -m pytest tests/test_vae_large_gpu.py -x -q > gpurun_out/p12_vae.txt 2>&1; tail -2 gpurun_out/p12_vae.txt
B=4096 timeout 100 python scripts/step_profile.py > gpurun_out/p12_step_B4096.txt 2>&1; head -4 gpurun_out/p12_step_B4096.txt; grep "conv2.wgrad" gpurun_out/p12_step_B4096.txt
B=512 timeout 100 python scripts/step_profile.py > gpurun_out/p12_step_B512.txt 2>&1; head -4 gpurun_out/p12_step_B512.txt; grep "conv2.wgrad" gpurun_out/p12_step_B512.txt
CPB_TC_DEBUG=16 timeout 100 python scripts/tc_prof.py > gpurun_out/p12_tcprof.txt 2>&1; grep tc3prof gpurun_out/p12_tcprof.txt | head -6
