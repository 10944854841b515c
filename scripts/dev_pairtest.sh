mkdir -p gpurun_out
B=128 timeout 100 python scripts/diag_dp.py > gpurun_out/dd1.txt 2>&1; grep -v Warn gpurun_out/dd1.txt | tail -23 | cut -c1-170
CPB_TC_PAIR=0 B=128 timeout 100 python scripts/diag_dp.py > gpurun_out/dd2.txt 2>&1; grep -v Warn gpurun_out/dd2.txt | tail -23 | cut -c1-110
