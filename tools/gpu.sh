#!/bin/bash
# tools/gpu.sh <task> [args] — the commands run on the B200 box (through `gpurun -- 'bash tools/gpu.sh <task>'`).
# Outputs land in gpurun_out/ (scratch); what is worth keeping is condensed with tools/ncu_summary.py into profiles/.
#   tests                 pytest -m gpu + smoke
#   bench                 bench.py (both arms)
#   launches              ncu launch list of the bench command (gpu__time_duration per launch)
#   ncu <kernel-regex> [skip] [bench args...]   one `ncu --set full` capture of a kernel of bench.py, exported as raw + source CSV
#   sweep                 tools/sweep.py (rgb_to_lab sizes, K sweep, S-CIELAB stage, full searches)
#   multi N               NCCL parity test + bench at 1 and N GPUs   (gpurun --gpus N)
#   micro                 FP32-pipe / issue-model microbenchmarks
#   final                 tests + bench + launches + latency in one call
#   cbench N              tools/multi_c_bench.c: single-process C host on 1..N devices (headline step + default search), both exchange paths
#   latency               tools/latency_ab.py: us per search iteration with / without the direct host I/O path
set -u
mkdir -p gpurun_out
task=${1:-tests}; shift || true
case "$task" in
tests)
    timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
bench)
    nproc > gpurun_out/host.txt; lscpu | grep "Model name" >> gpurun_out/host.txt
    timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
    timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
    tail -c 1500 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err ;;
launches)
    BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
    $BENCH > gpurun_out/plain.log 2>&1 &&
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
    echo "launch list rc=$?" ;;
ncu)
    regex=${1:?kernel regex}; skip=${2:-0}; shift; shift || true
    BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweeps $*"
    $BENCH > gpurun_out/plain_ncu.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/prof $BENCH > gpurun_out/ncu.log 2>&1
    echo "capture rc=$?"
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > gpurun_out/prof_source.csv 2>/dev/null
    # the summary kept under profiles/ and, for the headline kernel at the bench's default shape, the measured DRAM bytes per launch
    # (profiles/traffic.json = bench.py's roofline.traffic, tied to the sha of the kernel source: copy both back from gpurun_out/)
    if [ "$regex" = "assign_reduce_kernel" ] && [ -z "$*" ]; then
        python tools/ncu_summary.py gpurun_out/prof_raw.csv gpurun_out/prof_source.csv --units 530841600 \
            --traffic-key assign_reduce_w3840_h2160_k256_b64 --capture-name "profiles/r02/ncu_assign_reduce_final_4k_k256_b64.json" > gpurun_out/ncu_summary.json
        cp profiles/traffic.json gpurun_out/traffic.json
    else
        python tools/ncu_summary.py gpurun_out/prof_raw.csv gpurun_out/prof_source.csv > gpurun_out/ncu_summary.json
    fi
    ls -la gpurun_out/prof* ;;
sweep)
    timeout 1200 python tools/sweep.py "$@" > gpurun_out/sweep.json 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err; ls -la gpurun_out/sweep.json ;;
multi)
    N=${1:-2}
    nvidia-smi -L | head -8
    timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_native.py -x -q -m gpu 2>&1 | tail -6
    for n in 1 $N; do
        if [ "$n" -eq 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-sweeps > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
        else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$n" > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; fi
        python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n$n.json') if l.startswith('{')][-1])
    print('N=$n exhaustive', round(d['value'], 2), 'Gpixel/s', round(d['ms_per_step'], 3), 'ms/step, frac', round(d['roofline']['frac'], 3), '| e2e', round(d['e2e']['value'], 2), '| pruned', round(d['pruned']['value'], 1), 'Gpixel/s | clocks', d['clocks'])
    print('   parity', d.get('parity')); print('   strong_64mp', {k: v for k, v in (d.get('strong_64mp') or {}).items() if k in ('ms_per_step', 'value', 'clocks')})
except Exception as e:
    print('N=$n parse failed', e)
PY
    done ;;
micro)
    for m in microbench microbench2 microbench3 microbench4; do [ -x tools/$m ] && timeout 200 ./tools/$m > gpurun_out/$m.json 2> gpurun_out/$m.err; done; ls -la gpurun_out/microbench* ;;
scbench)
    timeout 300 python tools/sc_bench.py "$@" > gpurun_out/sc_bench.json 2> gpurun_out/sc_bench.err; tail -2 gpurun_out/sc_bench.err; cat gpurun_out/sc_bench.json
    timeout 300 python tools/sc_bench.py --once > gpurun_out/sc_plain.log 2>&1 &&
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/sc_launches.csv python tools/sc_bench.py --once > gpurun_out/sc_ncu.log 2>&1
    grep -E "sc_|assign|pruned" gpurun_out/sc_launches.csv | awk -F'","' '{print $5, $NF}' | tail -14 ;;
ncupy)
    # one `ncu --set full` capture of a kernel of any python command: ncupy <kernel-regex> <skip> <script and args...>
    regex=${1:?kernel regex}; skip=${2:-0}; shift; shift
    python "$@" > gpurun_out/plain_ncu.log 2>&1 &&
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/prof python "$@" > gpurun_out/ncu.log 2>&1
    echo "capture rc=$?"
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > gpurun_out/prof_source.csv 2>/dev/null
    ls -la gpurun_out/prof* ;;
final)
    # the evidence set of a finished state: GPU tests + smoke, both bench arms, the launch list of the bench command, latencies
    bash tools/gpu.sh tests
    bash tools/gpu.sh bench
    bash tools/gpu.sh launches
    bash tools/gpu.sh latency ;;
cbench)
    # tools/multi_c_bench.c: a plain C host on 1, 2, ... N devices of one process (gpurun --gpus N); once per exchange path
    N=${1:-2}
    gcc -O2 -Wall -Iinclude -o tools/multi_c_bench tools/multi_c_bench.c -Lhybridquantization_b200 -lhq_b200 -Wl,-rpath,$PWD/hybridquantization_b200 || exit 1
    for peers in 1 0; do HQ_PEER_EXCHANGE=$peers timeout 600 ./tools/multi_c_bench "$N" | tee -a gpurun_out/multi_c_bench_n$N.jsonl; done ;;
latency)
    timeout 500 python tools/latency_ab.py > gpurun_out/latency_ab.json 2> gpurun_out/latency_ab.err; tail -3 gpurun_out/latency_ab.err; ls -la gpurun_out/latency_ab.json ;;
*) echo "unknown task $task"; exit 2 ;;
esac
