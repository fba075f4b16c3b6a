#!/bin/bash
mkdir -p gpurun_out
echo "== microbench2"; timeout 120 ./tools/microbench2 > gpurun_out/microbench2.json 2>&1
echo "== pytest gpu (parity + swasa)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_swasa.py -x -q -m gpu 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'])"
tail -3 gpurun_out/bench.err
