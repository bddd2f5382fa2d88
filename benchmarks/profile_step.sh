#!/bin/bash
# ncu evidence for the step kernel (run on the GPU box from the repo root, after `python bench.py` exited 0):
#   1. --set full capture of one k_rollout_tab launch at the bench workload -> gpurun_out/prof_r2_tab.ncu-rep
#   2. launch list of the bench command (per-launch gpu__time_duration)     -> gpurun_out/r2_launches.csv
set -x
mkdir -p gpurun_out
python benchmarks/step_profile_target.py 1 > gpurun_out/step_target.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rollout_tab -s 2 -c 1 -f \
    -o gpurun_out/prof_r2_tab python benchmarks/step_profile_target.py 1 > gpurun_out/ncu_tab.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
