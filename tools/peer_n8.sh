#!/bin/bash
# tools/peer_n8.sh N — multi-GPU parity tests (torchrun ranks + single-process C host at 2..N devices) and bench.py at N ranks
set -u
N=${1:-8}
mkdir -p gpurun_out
export HQ_PEER_TIMEOUT_MS=20000
nvidia-smi -L | head -8
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_native.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_multi_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus "$N" --no-cpu-baseline --force-sweeps > gpurun_out/bench_n${N}_peer.json 2> gpurun_out/bench_n${N}_peer.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n${N}_peer.err
python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n${N}_peer.json') if l.startswith('{')][-1])
    print('value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'frac', round(d['roofline']['frac'], 3), 'parity', d.get('parity'), d['clocks'])
    print('strong', {k: v for k, v in d['strong_64mp'].items() if k in ('ms_per_step', 'value', 'clocks')})
    for row in d['k_sweep']['rows']:
        print('  K', row['K'], {k: (round(v['ms_per_step'] * 1e3, 1), round(v['gpixel_per_s'], 1), round(v['frac'], 3)) for k, v in row.items() if isinstance(v, dict)})
except Exception as e:
    print('parse failed', e)
PY
