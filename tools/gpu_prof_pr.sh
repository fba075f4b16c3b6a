#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain_pr.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:pruned_assign -s 2 -c 1 -o gpurun_out/prof_pr $BENCH > gpurun_out/ncu_pr.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/prof_pr.ncu-rep --page raw --csv > gpurun_out/prof_pr_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_pr.ncu-rep --page source --csv > gpurun_out/prof_pr_source.csv 2>/dev/null
ls -la gpurun_out/prof_pr*
