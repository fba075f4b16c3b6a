#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --rows-per-gpu 8640"
$BENCH > gpurun_out/plain_rl.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rgb_to_lab -c 1 -o gpurun_out/prof_rl $BENCH > gpurun_out/ncu_rl.log 2>&1
echo "rc=$?"; ls -la gpurun_out/prof_rl.ncu-rep
