#!/bin/bash
# bench + ncu full capture of the assign kernel + launch list (bounded)
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$BENCH > gpurun_out/plain2.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:assign_reduce -s 3 -c 1 -o gpurun_out/prof_assign $BENCH > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out | tail -8
