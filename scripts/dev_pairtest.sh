mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_vae_gpu.py tests/test_vae_large_gpu.py -x -q > gpurun_out/p13_vae.txt 2>&1; tail -4 gpurun_out/p13_vae.txt
B=4096 timeout 100 python scripts/step_profile.py > gpurun_out/p13_step_B4096.txt 2>&1; head -36 gpurun_out/p13_step_B4096.txt
B=512 timeout 100 python scripts/step_profile.py > gpurun_out/p13_step_B512.txt 2>&1; head -3 gpurun_out/p13_step_B512.txt
