#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01J.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_pcg_persistent|k_symv_lower|k_cluster_inverse' --launch-skip 20 --launch-count 4 -o gpurun_out/prof_r01J python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/prof_r01J.ncu-rep
