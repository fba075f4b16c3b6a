#!/bin/bash
# multi-GPU: NCCL parity test + bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== multi-gpu test"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
for n in 1 $N; do
  echo "== bench N=$n"
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n1.err > gpurun_out/bench_n1.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/bench_n$n.err > gpurun_out/bench_n$n.json
  fi
  tail -2 gpurun_out/bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$n.json') if l.startswith('{')][-1])
    print('N=$n value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
except Exception as e: print('parse failed', e)
PY
done
