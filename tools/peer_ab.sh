#!/bin/bash
# tools/peer_ab.sh N — the exchange over NVLink peer memory against ncclAllReduce on N GPUs (gpurun --gpus N):
# the multi-GPU parity tests, then bench.py at N ranks with the K sweep forced, once per exchange path.
set -u
N=${1:-2}
mkdir -p gpurun_out
export HQ_PEER_TIMEOUT_MS=20000
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_native.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_multi_n$N.txt
HQ_SOAK_ITERS=20000 HQ_SOAK_LAUNCHES=5000 timeout 900 python -m pytest tests/test_gpu_multi_soak.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_soak_n$N.txt
for peers in 1 0; do
    HQ_PEER_EXCHANGE=$peers timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 2951$peers \
        bench.py --gpus "$N" --no-cpu-baseline --force-sweeps > gpurun_out/bench_n${N}_peers$peers.json 2> gpurun_out/bench_n${N}_peers$peers.err
    echo "bench peers=$peers rc=$?"; tail -3 gpurun_out/bench_n${N}_peers$peers.err
    python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n${N}_peers$peers.json') if l.startswith('{')][-1])
    print('peers=$peers value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'parity', d.get('parity', {}).get('ok'), d.get('parity', {}).get('peer_exchange'), d['clocks'])
    for row in d['k_sweep']['rows']:
        print('  K', row['K'], {k: (round(v['ms_per_step'] * 1e3, 1), round(v['gpixel_per_s'], 1)) for k, v in row.items() if isinstance(v, dict)})
except Exception as e:
    print('parse failed', e)
PY
done
