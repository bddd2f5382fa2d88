#!/usr/bin/env python
"""profiles/r2_traffic.json from an `ncu --set full` capture of the step kernel: DRAM bytes per env-step, stamped with
the hash of the kernel sources the capture was taken from (bench.py refuses the file when the sources have changed).
Usage: python profiles/make_traffic.py gpurun_out/prof_r2_tab.ncu-rep <env-steps per launch>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

rep, env_steps = sys.argv[1], float(sys.argv[2])
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                                                  text=True).stdout)))
h, units, row = raw[0], raw[1], raw[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def val(name):
    i = h.index(name)
    return float(row[i]) * scale[units[i]]


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {"kernel": row[h.index("Kernel Name")][:60], "report": os.path.basename(rep), "env_steps_per_launch": env_steps,
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_env_step": (rd + wr) / env_steps,
       "algorithmic_bytes_per_env_step": 79.0, "gpu_time_ms": float(row[h.index("gpu__time_duration.sum")]),
       "source_sha256": kernel_source_hash()}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(out)
