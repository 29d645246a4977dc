#!/usr/bin/env python
"""Dev helper: per-function SASS opcode histogram, and the histogram of the hot loop
(the backward-branch body with the most FFMA/FFMA2).  usage: tools_sass.py lib.so substring"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("==", name, len(ins), "instructions")
    # find loops: BRA to lower address
    best = None
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`\(\.L_x_\d+\)|BRA\S*\s+0x([0-9a-f]+)", t)
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t)
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt < a:
                    body = [x for x in ins if tgt <= x[0] <= a]
                    nf = sum(1 for x in body if "FFMA" in x[1] or "FMUL" in x[1] or "FADD" in x[1])
                    # innermost loop with real FP work: smallest body holding >= 40 FP instructions
                    if nf >= 40 and (best is None or len(body) < len(best[3])):
                        best = (nf, tgt, a, body)
    def hist(body):
        c = collections.Counter()
        for _, t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t.split()[0].split(".")[0]] += 1
        return c
    if best:
        nf, tgt, a, body = best
        print(f"-- hot loop 0x{tgt:x}..0x{a:x}: {len(body)} instructions")
        for k, v in hist(body).most_common():
            print(f"   {v:5d} {k}")
    else:
        for k, v in hist(ins).most_common(25):
            print(f"   {v:5d} {k}")
