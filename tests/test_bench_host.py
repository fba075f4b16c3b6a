"""bench.py's host-side contract, checked without a GPU: the reference arm (`--impl reference`, the reference's own kernels
compiled for the CPU — or the oracle port where oracle/_ref is absent) prints ONE JSON line with the keys the driver reads, and the
clock sampler degrades to "unavailable" instead of failing where neither NVML nor nvidia-smi exists."""
import importlib.util
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("hq_bench", os.path.join(REPO, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_reference_arm_prints_the_contract_line(oracle):
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cielab_assign_gpixel_per_s_4k_k256" and d["unit"] == "Gpixel/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "3840x2160" in d["config"]["workload"]


def test_clock_sampler_degrades_without_a_gpu():
    m = _bench_module()
    s = m.ClockSampler(0)
    s.start()
    ready = s.wait_ready(0.5)
    s.stop()
    out = s.summary(0.0, 1e18)
    if not ready:   # this container: no NVML device, no nvidia-smi
        assert out == {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    else:
        assert out["samples"] >= 1 and out["source"] in ("nvml", "nvidia-smi")
