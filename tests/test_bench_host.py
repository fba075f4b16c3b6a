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


def test_traffic_figure_is_dropped_when_the_kernel_source_changed(tmp_path, monkeypatch):
    """roofline.traffic comes from the last ncu capture (profiles/traffic.json, written by tools/ncu_summary.py --traffic-key with
    the sha of the kernel source): bench.py reports it only while that source is unchanged."""
    m = _bench_module()
    repo = tmp_path / "repo"
    (repo / "profiles").mkdir(parents=True)
    src = repo / "hybridquantization_b200" / "csrc"
    src.mkdir(parents=True)
    (src / "hq_kernels.cu").write_text("// kernel v1\n")
    monkeypatch.setattr(m, "REPO", str(repo))
    json.dump({"k": {"bytes": 123, "capture": "cap.json", "source_file": "hq_kernels.cu", "source_sha": m._file_sha(str(src / "hq_kernels.cu"))},
               "old_style": 7, "source": "legacy"}, open(repo / "profiles" / "traffic.json", "w"))
    assert m.measured_traffic("k") == (123, "cap.json")
    assert m.measured_traffic("old_style") == (7, "legacy")
    assert m.measured_traffic("absent")[0] is None
    (src / "hq_kernels.cu").write_text("// kernel v2\n")
    v, why = m.measured_traffic("k")
    assert v is None and "stale" in why


def test_ncu_summary_writes_the_traffic_entry(tmp_path):
    """tools/ncu_summary.py --traffic-key: dram read + write bytes of the captured launch, keyed, with the source sha."""
    raw = tmp_path / "raw.csv"
    raw.write_text('"ID","Kernel Name","dram__bytes_read.sum","dram__bytes_write.sum","gpu__time_duration.sum"\n'
                   '"","","Mbyte","Kbyte","us"\n"0","k<3>","99.5","6300.0","19.7"\n')
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    before = open(tpath).read()
    try:
        r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_summary.py"), str(raw), "--traffic-key", "_unit_test_key",
                            "--capture-name", "unit.json"], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        e = json.load(open(tpath))["_unit_test_key"]
        assert e["bytes"] == 99500000 + 6300000 and e["capture"] == "unit.json" and e["source_file"] == "hq_kernels.cu" and len(e["source_sha"]) == 12
    finally:
        open(tpath, "w").write(before)
