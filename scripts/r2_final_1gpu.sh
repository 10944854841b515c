#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): tests, smoke, bench (both arms), per-group profiles, cycle log.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r2_pytest_gpu.txt
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
timeout 150 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; head -c 300 gpurun_out/r2_bench_reference_arm.json; echo
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; head -c 400 gpurun_out/r2_bench_1gpu.json; echo
B=4096 OUT=gpurun_out/r2_groups_B4096.json timeout 150 python scripts/step_profile.py > gpurun_out/r2_groups_B4096.txt 2>&1; head -3 gpurun_out/r2_groups_B4096.txt
B=512 OUT=gpurun_out/r2_groups_B512.json timeout 150 python scripts/step_profile.py > gpurun_out/r2_groups_B512.txt 2>&1; head -3 gpurun_out/r2_groups_B512.txt
CPB_TC_DEBUG=16 timeout 150 python scripts/tc_prof.py > gpurun_out/r2_tcprof_final.txt 2>&1; grep -c tc2prof gpurun_out/r2_tcprof_final.txt
timeout 150 python bench.py --config 5 > gpurun_out/r2_config5_1gpu.json 2> gpurun_out/r2_config5_1gpu.err; cat gpurun_out/r2_config5_1gpu.json
