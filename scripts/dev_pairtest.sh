mkdir -p gpurun_out
export WEIGHTS=shipped
FRAMES=shipped SEED=0 B=4 timeout 100 python scripts/diag_smoke.py > gpurun_out/d5.txt 2>&1; grep -v Warning gpurun_out/d5.txt | tail -23 | cut -c1-80
CPB_TC_PAIR=0 FRAMES=shipped SEED=0 B=4 timeout 100 python scripts/diag_smoke.py > gpurun_out/d6.txt 2>&1; tail -23 gpurun_out/d6.txt | cut -c1-80
SEED=0 B=4 timeout 100 python scripts/diag_smoke.py > gpurun_out/d7.txt 2>&1; tail -23 gpurun_out/d7.txt | cut -c1-80 | head -5
FRAMES=shipped SEED=8 B=8 timeout 100 python scripts/diag_smoke.py > gpurun_out/d8.txt 2>&1; tail -23 gpurun_out/d8.txt | cut -c1-80 | head -5
