#!/bin/bash
# first GPU pass: smoke, FP32 microbenchmark, parity tests, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep "Model name" >> gpurun_out/host.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== microbench"; timeout 120 ./tools/microbench > gpurun_out/microbench.json 2> gpurun_out/microbench.err; tail -c 600 gpurun_out/microbench.json; echo
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
