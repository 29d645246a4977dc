#!/usr/bin/env python
"""Fold fresh ncu exports of single modes into profiles/r2_ncu_ring_modes.json and profiles/r2_ncu_source_segments.txt.
On the GPU box, per mode M (tools/profile_modes.py):
  python tools/profile_modes.py M && ncu --set full --clock-control none --import-source on -k regex:ring_step_kernel -s 3 -c 1 \
      -f -o /tmp/M python tools/profile_modes.py M && ncu -i /tmp/M.ncu-rep --page raw --csv > gpurun_out/ncu_M_raw.csv && \
      ncu -i /tmp/M.ncu-rep --page source --csv > gpurun_out/ncu_M_src.csv
then here:  refresh_mode_profiles.py M [M ...]"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JS = os.path.join(ROOT, "profiles", "r2_ncu_ring_modes.json")
SEG = os.path.join(ROOT, "profiles", "r2_ncu_source_segments.txt")


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[h], rows[h + 1]
    v = next(r for r in rows[h + 2:] if r and r[0].isdigit())
    out = {}
    for n, u, x in zip(names, units, v):
        out[n] = f"{x} {u}".strip() if u else x
    return out


js = json.load(open(JS))
keep = list(next(iter(js.values())).keys())
seg_text = open(SEG).read()
for mode in sys.argv[1:]:
    m = raw_metrics(os.path.join(ROOT, "gpurun_out", f"ncu_{mode}_raw.csv"))
    js[mode] = {k: m[k] for k in keep if k in m}
    e = js[mode]
    head = (f"==== {mode}:  {e['Kernel Name']}\n     duration {e['gpu__time_duration.sum']}, {e['launch__registers_per_thread']}, "
            f"warp instructions {e['smsp__inst_executed.sum']}, FMA pipe active "
            f"{e['sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']}, issue active "
            f"{e['smsp__issue_active.avg.pct_of_peak_sustained_active']}, DRAM read {e['dram__bytes_read.sum']} write "
            f"{e['dram__bytes_write.sum']}\n")
    body = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_segments.py"),
                           os.path.join(ROOT, "gpurun_out", f"ncu_{mode}_src.csv")], capture_output=True, text=True).stdout
    block = head + body + "\n"
    pat = re.compile(r"==== " + re.escape(mode) + r":.*?(?=\n==== |\Z)", re.S)
    if pat.search(seg_text):
        seg_text = pat.sub(lambda _: block.rstrip("\n") + "\n", seg_text, count=1)
    else:
        seg_text = seg_text.rstrip("\n") + "\n\n" + block
    print(mode, e["gpu__time_duration.sum"], e["smsp__inst_executed.sum"])
json.dump(js, open(JS, "w"), indent=1)
open(SEG, "w").write(seg_text)
