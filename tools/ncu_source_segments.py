#!/usr/bin/env python
"""Group the SASS rows of an `ncu --page source --csv` export into segments of equal execution count and print, per
segment: address range, instructions, executions per instruction, share of all executed warp-instructions, share of the
stall samples, and the dominant opcodes.  usage: ncu_source_segments.py source.csv [min_share_percent]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
cols = rows[h]
ix = {c: i for i, c in enumerate(cols)}
data = [r for r in rows[h + 1:] if len(r) > ix["Instructions Executed"] and r[ix["Instructions Executed"]].strip() != ""]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]] or 0) for r in data)
print(f"total warp-instructions executed {tot_inst}, stall samples {tot_samp}, SASS instructions {len(data)}")
segs = []
cur = None
for r in data:
    n = int(r[ix["Instructions Executed"]])
    # same segment if the count is within 2 % of the segment's first count
    if cur is not None and (n == cur["n0"] or (cur["n0"] > 0 and abs(n - cur["n0"]) <= 0.02 * cur["n0"])):
        cur["rows"].append(r)
    else:
        cur = {"n0": n, "rows": [r]}
        segs.append(cur)
stall_cols = [c for c in cols if c.startswith("stall_") and "Not Issued" not in c]
for s in segs:
    inst = sum(int(r[ix["Instructions Executed"]]) for r in s["rows"])
    samp = sum(int(r[ix["# Samples"]] or 0) for r in s["rows"])
    if 100.0 * inst / tot_inst < min_share and 100.0 * samp / max(1, tot_samp) < min_share:
        continue
    ops = collections.Counter(r[ix["Source"]].split()[0].split(".")[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1].split(".")[0] for r in s["rows"])
    st = collections.Counter()
    for r in s["rows"]:
        for c in stall_cols:
            v = r[ix[c]]
            if v:
                st[c] += int(v)
    top_st = ", ".join(f"{k[6:]} {100.0 * v / max(1, samp):.0f}%" for k, v in st.most_common(4))
    print(f"{s['rows'][0][ix['Address']][-6:]}..{s['rows'][-1][ix['Address']][-6:]}  {len(s['rows']):4d} instr x {s['n0']:9d}  "
          f"{100.0 * inst / tot_inst:5.1f}% of instr  {100.0 * samp / max(1, tot_samp):5.1f}% of samples  [{top_st}]  "
          + " ".join(f"{k}:{v}" for k, v in ops.most_common(7)))
