#!/usr/bin/env python
"""SASS evidence: per kernel of ppo_car_b200/libcarenv_b200.so, how many of the mnemonics that prove the
Blackwell-native paths (tcgen05 = UTCHMMA / UTCBAR / LDTM / UTCATOMSWS, packed f32x2 = FFMA2 / FMUL2, 3-input
min/max = FMNMX3, warp reduce = CREDUX) the code contains.  Usage: python profiles/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ppo_car_b200", "libcarenv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "FFMA2", "FMUL2", "FADD2", "FMNMX3", "CREDUX", "FMUL.SAT", "DFMA", "LDS", "LDCU",
         "MUFU", "HMMA", "UTMALDG"]
per, cur, arch = collections.OrderedDict(), None, set()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|carenv::", "", name).split("(")[0]
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w == "FMUL.SAT" and op.startswith("FMUL.SAT")):
                cur[w] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(per)} kernels, arch {sorted(arch)}")
print("| kernel | instr | " + " | ".join(WATCH) + " |")
print("|---|---|" + "---|" * len(WATCH))
for k, c in per.items():
    if c["total"] < 200:
        continue
    print(f"| `{k[:70]}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")
