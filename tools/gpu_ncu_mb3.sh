#!/bin/bash
mkdir -p gpurun_out
for vr in "0 4" "1 8" "3 4"; do
  set -- $vr
  timeout 60 ./tools/microbench3 $1 $2 > gpurun_out/mb3_$1_$2.json 2>&1 &&
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:sweep -s 1 -c 1 -o gpurun_out/prof_mb3_v$1_r$2 ./tools/microbench3 $1 $2 > gpurun_out/ncu_mb3_$1_$2.log 2>&1
  echo "V$1 R$2 rc=$?"
done
ls -la gpurun_out
