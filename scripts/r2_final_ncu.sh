#!/bin/bash
# Round-2 ncu evidence on ONE B200: launch list of the bench command + one --set full capture of the tensor-core kernels.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
wc -l gpurun_out/r2_launches.csv
# second train step of scripts/tc_prof.py: the 12 tap-GEMM + 6 weight-gradient launches (skip the first step's 18)
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"tc2_tapgemm_kernel|tc_wgrad_kernel" --launch-skip 18 -c 18 \
    -o gpurun_out/r2_full_tc -f python scripts/tc_prof.py > gpurun_out/r2_ncu_full_tc.log 2>&1
ncu -i gpurun_out/r2_full_tc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw_tc.csv 2>/dev/null; wc -l gpurun_out/r2_ncu_full_raw_tc.csv
rm -f gpurun_out/r2_full_tc.ncu-rep     # gpurun merges at most 64 MiB back: keep the CSV pages only
timeout 400 ncu --set full --clock-control none -k regex:"edge_|deconv4_fwd|recon_loss|colsum_kernel|prep_frames|adam_kernel" --launch-skip 18 -c 22 \
    -o gpurun_out/r2_full_other -f python scripts/tc_prof.py > gpurun_out/r2_ncu_full_other.log 2>&1
ncu -i gpurun_out/r2_full_other.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw_other.csv 2>/dev/null; wc -l gpurun_out/r2_ncu_full_raw_other.csv
rm -f gpurun_out/r2_full_other.ncu-rep
