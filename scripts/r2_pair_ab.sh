#!/bin/bash
# A/B of the CTA-pair tap-GEMM (default) against 1-CTA MMAs on the final build: per-group step times at both batch sizes.
mkdir -p gpurun_out
for b in 4096 512; do
  CPB_TC_PAIR=0 B=$b timeout 120 python scripts/step_profile.py > gpurun_out/r2_groups_pair0_B$b.txt 2>&1; head -1 gpurun_out/r2_groups_pair0_B$b.txt
  CPB_TC_PAIR=1 B=$b timeout 120 python scripts/step_profile.py > gpurun_out/r2_groups_pair1_B$b.txt 2>&1; head -1 gpurun_out/r2_groups_pair1_B$b.txt
done
