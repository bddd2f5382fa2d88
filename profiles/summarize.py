#!/usr/bin/env python
"""Turn an ncu report (captured with `--set full --import-source on`) into the text summaries kept
under profiles/: headline metrics, executed-instruction mix per env-step, and stall samples per
source line.  Usage: python profiles/summarize.py gpurun_out/prof.ncu-rep <env-steps per launch> > profiles/x.txt"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, env_steps = sys.argv[1], float(sys.argv[2])
kern = sys.argv[3] if len(sys.argv) > 3 else "rolloutIhhLi4"
W = env_steps / 32.0


def ncu(page):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True,
                                                      text=True).stdout)))


raw = ncu("raw")
h = raw[0]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second"]
KEYS += [k for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")]
print(f"# {os.path.basename(rep)}  ({raw[2][h.index('Kernel Name')][:80]})")
print("## metrics (first captured launch)")
for k in KEYS:
    if k in h:
        print(f"{k} [{raw[1][h.index(k)]}] = {raw[2][h.index(k)]}")

src = ncu("source")
hdr = [r for r in src if r and r[0] == "Address"][0]
body = [r for r in src if r and re.match(r"^[0-9a-fx]+$", r[0].strip()) and len(r) == len(hdr)]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(float(r[ie] or 0) for r in body)
tsamp = sum(float(r[isamp] or 0) for r in body)
op, samp = collections.Counter(), collections.Counter()
for r in body:
    m = re.match(r"^(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", r[ia].strip())
    if m:
        op[m.group(2)] += float(r[ie] or 0)
        samp[m.group(2)] += float(r[isamp] or 0)
print(f"\n## executed warp-instructions per env-step (per thread): total {tot / W:.1f}")
for k, v in op.most_common(28):
    print(f"{k:10s} {v / W:8.1f}  {100 * v / tot:5.1f}% of instructions  {100 * samp[k] / tsamp:5.1f}% of stall samples")

# source-line attribution through nvdisasm line info of the in-tree library
lib = os.path.join(ROOT, "ppo_car_b200", "libcarenv_b200.so")
tmp = "/tmp/_carenv_prof"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
if cub:
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub[0])], capture_output=True, text=True).stdout
    cur, inside, a2l = None, False, {}
    for ln in dis.splitlines():
        if ".text." in ln:
            inside = kern in ln
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            a2l[int(m.group(1), 16)] = cur
    base = int(body[0][0], 16)
    agg = collections.defaultdict(lambda: [0.0, 0.0])
    for r in body:
        k = a2l.get(int(r[0], 16) - base)
        agg[k][0] += float(r[isamp] or 0)
        agg[k][1] += float(r[ie] or 0) / W
    files = {}

    def text(k):
        if not k:
            return ""
        p = os.path.join(ROOT, "ppo_car_b200", "csrc", k[0])
        if os.path.exists(p):
            files.setdefault(p, open(p).read().splitlines())
            return files[p][k[1] - 1].strip()[:100]
        return "(CUDA header: packed f32x2 intrinsics)" if "sm_100_rt" in k[0] else ""

    print("\n## stall samples by source line (top 30)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f"{100 * v[0] / tsamp:5.1f}%  {v[1]:7.1f} inst/env-step  {k}  {text(k)}")
