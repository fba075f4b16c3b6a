#!/bin/bash
# final r01 evidence: bench (full), launch list, ncu full of assign + S-CIELAB kernels
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -c 400 gpurun_out/bench.json; echo
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$BENCH > gpurun_out/plain2.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:assign_reduce -s 3 -c 1 -o gpurun_out/prof_assign $BENCH > gpurun_out/ncu_full.log 2>&1
echo "assign capture rc=$?"
SC="python tools/sweep.py --quick --skip-swasa --only-scielab"
$SC > gpurun_out/sc_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sc_.pass21 -s 4 -c 2 -o gpurun_out/prof_scielab $SC > gpurun_out/ncu_sc.log 2>&1
echo "scielab capture rc=$?"; ls -la gpurun_out | tail -8
