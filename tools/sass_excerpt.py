#!/usr/bin/env python
"""SASS excerpts of the hot loops for profiles/: `cuobjdump -sass` of the built objects, cut down to the loop of a kernel
that holds the most floating-point instructions, with its instruction histogram on top.
  python tools/sass_excerpt.py <out-dir>"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(REPO, "hybridquantization_b200", "build")
KERNELS = [  # (object, mangled-name fragment, output name, what it is)
    ("hq_kernels.o", "assign_reduce_kernelILi3ELb0ELb0ELi0E", "assign_reduce_v3", "assign_reduce_kernel<3,false,false,0>: prefilter sweep, K > 32 (bench headline)"),
    ("hq_kernels.o", "assign_reduce_kernelILi1ELb0ELb0ELi0E", "assign_reduce_v1", "assign_reduce_kernel<1,false,false,0>: direct form, K <= 32"),
    ("hq_kernels.o", "rgb_to_lab_kernelILb0E", "rgb_to_lab", "rgb_to_lab_kernel<false>: packed f32x2 pixel pairs"),
    ("hq_pruned.o", "pruned_assign_kernel", "pruned_assign", "pruned_assign_kernel: exact sweep over the surviving colours"),
    ("hq_scielab.o", "sc_candidate_strip21_kernelIhE", "sc_strip21", "sc_candidate_strip21_kernel<u8>: S-CIELAB candidate stage"),
]
PAT = re.compile(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);")


def op(text):
    t = text.split()
    o = t[1] if t[0].startswith("@") else t[0]
    return o.split(".")[0]


def main():
    out = sys.argv[1]
    os.makedirs(out, exist_ok=True)
    cache = {}
    for obj, frag, name, what in KERNELS:
        if obj not in cache:
            cache[obj] = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
        txt = cache[obj]
        i = txt.find(frag)
        if i < 0:
            print("not found:", frag)
            continue
        j = txt.find("Function :", i + 10)
        ins = [(int(m.group(1), 16), m.group(2).strip()) for m in PAT.finditer(txt[i:j if j > 0 else len(txt)])]
        best = None
        for a, t in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                seg = [(x, y) for x, y in ins if int(m.group(1), 16) <= x <= a]
                fp = sum(op(y) in ("FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2") for _, y in seg)
                score = fp / len(seg) if len(seg) >= 40 else 0.0   # the densest loop, not the outermost one
                if best is None or score > best[0]:
                    best = (score, seg)
        seg = best[1]
        hist = collections.Counter(op(y) for _, y in seg)
        with open(os.path.join(out, f"sass_{name}_hot_loop.txt"), "w") as f:
            f.write(f"// {what}\n// cuobjdump -sass {obj} (sm_100a, nvcc 12.9), hot loop 0x{seg[0][0]:x}..0x{seg[-1][0]:x}: {len(seg)} instructions\n")
            f.write("// " + ", ".join(f"{k} {v}" for k, v in hist.most_common()) + "\n")
            for a, t in seg:
                f.write(f"/*{a:05x}*/ {t} ;\n")
        print(name, len(seg), hist.most_common(6))


if __name__ == "__main__":
    main()
