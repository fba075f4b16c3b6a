#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --colors ${1:-8} --batch ${2:-1}"
$BENCH > gpurun_out/plain_sk.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:assign_reduce -s 3 -c 1 -o gpurun_out/prof_sk $BENCH > gpurun_out/ncu_sk.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/prof_sk.ncu-rep --page raw --csv > gpurun_out/prof_sk_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_sk.ncu-rep --page source --csv > gpurun_out/prof_sk_source.csv 2>/dev/null
