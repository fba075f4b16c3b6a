#!/bin/bash
# microbench2 + bench + ncu launch list + ncu full capture of the assign kernel
mkdir -p gpurun_out
echo "== microbench2"; timeout 120 ./tools/microbench2 > gpurun_out/microbench2.json 2> gpurun_out/microbench2.err; tail -c 300 gpurun_out/microbench2.json; echo
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
$BENCH > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$BENCH > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assign_reduce -s 3 -c 1 -o gpurun_out/prof_assign $BENCH > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out | tail -20
