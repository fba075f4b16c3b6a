#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` (+ optionally `--page source --csv`) into the small JSON summaries
kept under profiles/: the metrics DESIGN.md quotes, the stall breakdown, and the executed-instruction mix per pixel.
  python tools/ncu_summary.py raw.csv [source.csv] --units N > profiles/rNN/ncu_<kernel>.json"""
import argparse
import collections
import csv
import json

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("source", nargs="?")
    ap.add_argument("--units", type=float, default=0, help="pixels (or pixel x candidate pairs) one launch processes")
    ap.add_argument("--row", type=int, default=0, help="which captured launch (0 = first)")
    ap.add_argument("--traffic-key", help="also record dram read + write bytes of this launch in profiles/traffic.json under this key "
                                          "(bench.py's roofline.traffic), with the sha of the kernel source it was captured on")
    ap.add_argument("--source-file", default="hq_kernels.cu", help="csrc file holding the captured kernel (for the staleness check)")
    ap.add_argument("--capture-name", default="", help="name of the summary file this capture is kept as")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + a.row]
    out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None, "metrics": {}, "stalls_per_issue": {}}
    for k in KEYS:
        if k in hdr:
            out["metrics"][k] = {"value": vals[hdr.index(k)], "unit": units[hdr.index(k)]}
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if v >= 0.05:
                out["stalls_per_issue"][h.split("stalled_")[1].split("_per")[0]] = round(v, 3)
    if a.source:
        src = list(csv.reader(open(a.source)))
        h2 = src[1]
        si, xi = h2.index("Source"), h2.index("Instructions Executed")
        mix, tot = collections.Counter(), 0
        for r in src[2:]:
            if len(r) <= xi or not r[xi].isdigit():
                continue
            op = r[si].split()
            if not op:
                continue
            o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
            mix[o] += int(r[xi]); tot += int(r[xi])
        out["warp_instructions_executed"] = tot
        if a.units:
            out["thread_instructions_per_unit"] = round(tot * 32 / a.units, 2)
            out["instruction_mix_per_unit"] = {o: round(n * 32 / a.units, 2) for o, n in mix.most_common(24)}
        else:
            out["instruction_mix"] = dict(mix.most_common(24))
    if a.traffic_key:
        import hashlib
        import os
        repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = out["metrics"][k]
            total += float(m["value"]) * to_bytes[m["unit"]]
        tpath = os.path.join(repo, "profiles", "traffic.json")
        t = json.load(open(tpath)) if os.path.exists(tpath) else {}
        src = os.path.join(repo, "hybridquantization_b200", "csrc", a.source_file)
        t[a.traffic_key] = {"bytes": int(round(total)), "capture": a.capture_name or "ncu --set full, one launch", "source_file": a.source_file,
                            "source_sha": hashlib.sha256(open(src, "rb").read()).hexdigest()[:12]}
        json.dump(t, open(tpath, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
