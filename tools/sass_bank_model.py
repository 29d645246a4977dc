#!/usr/bin/env python
"""Register-file model check.

    python tools/sass_bank_model.py <binary or .so> [--func SUBSTR] [--measured FILE]

For every kernel whose name contains SUBSTR (default: every kernel) find its hottest loop (the backward-branch body with
the most FP instructions) and, for every FFMA/FMUL/FADD/FFMA2/FMUL2/FADD2 in it, the source registers it reads from the
register file (immediates, constant-bank, uniform registers and RZ are free; an operand whose slot carried `.reuse` on the
previous FP instruction with the same register is served by the operand-reuse cache).  Then predict the issue cost per
instruction under three register-file hypotheses:

  port2   the sub-partition's register file delivers two 32-bit operands per lane per cycle, whatever their numbers:
          cost = max(base, reads / 2)                   (base = 1 scalar, 2 packed; a packed operand is 2 reads)
  bank2   two banks (reg % 2), each 64 bits wide: a conflict (one extra cycle) only when three scalar sources share a bank
  bank4   four single-ported banks (reg % 4): cost = max(base, most reads landing in one bank)

and print the predicted warp-instructions/clk/SMSP of the loop for each, next to the measured value when a
`--measured` file (the stdout of tools/fma_bank_microbench) is given.  The hypothesis whose predictions match ALL variants is
the register file; the same model is then applied to the dynamics kernel's substep loop (profiles/r2_sass_bank_model.txt)."""
from __future__ import annotations

import argparse
import collections
import math
import re
import subprocess

FP = ("FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD")


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        yield name, ins


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def hot_loop(ins, min_fp=40):
    """innermost loop with real FP work: the smallest backward-branch body holding >= min_fp FP instructions"""
    best = None
    for a, t in ins:
        if "BRA" not in t:
            continue
        m = re.search(r"0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        body = [x for x in ins if tgt <= x[0] <= a]
        nf = sum(1 for x in body if any(x[1].lstrip("@!UP0123456789 ").startswith(o) for o in FP))
        if nf >= min_fp and (best is None or len(body) < len(best[1])):
            best = (nf, body)
    return best[1] if best else []


def parse_fp(text):
    """-> (opcode, packed, [(reg, reuse)] register-file sources in slot order) or None"""
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    op = t.split()[0].split(".")[0]
    if op not in FP:
        return None
    packed = op.endswith("2")
    ops = [o.strip() for o in t.split(None, 1)[1].split(",")]
    srcs = []
    for o in ops[1:]:
        m = re.match(r"^[-|~]*R(\d+)((?:\.[A-Za-z0-9_]+)*)\|?$", o)
        if not m:
            srcs.append(None)      # immediate / c[][] / UR / RZ
            continue
        srcs.append((int(m.group(1)), ".reuse" in m.group(2)))
    return op, packed, srcs


def model(body):
    cost = collections.Counter()
    n = 0
    hist = collections.Counter()
    prev = None
    for _, t in body:
        p = parse_fp(t)
        if p is None:
            if not t.startswith(("NOP", "BRA", "ISETP", "IADD", "UIADD", "UISETP")):
                prev = None if not t.split()[0].startswith(("MOV", "IMAD")) else prev
            continue
        op, packed, srcs = p
        n += 1
        base = 2 if packed else 1
        reads = []          # 32-bit register reads that go to the register file
        for slot, s in enumerate(srcs):
            if s is None:
                continue
            reg, _ = s
            if prev is not None and slot < len(prev) and prev[slot] is not None and prev[slot] == (reg, True):
                continue    # operand-reuse cache hit
            reads += [reg, reg + 1] if packed else [reg]
        reads = sorted(set(reads))
        hist[(op, len(reads))] += 1
        cost["port2"] += max(base, len(reads) / 2)     # operand collector is pipelined: 1.5 cycles for 3 reads
        b2 = collections.Counter(r % 2 for r in reads)
        cost["bank2"] += max(base, base + (1 if (b2 and max(b2.values()) >= 3 and not packed) else 0),
                             math.ceil(max(b2.values()) / 2) if (packed and b2) else 0)
        b4 = collections.Counter(r % 4 for r in reads)
        cost["bank4"] += max(base, max(b4.values()) if b4 else 0)
        prev = srcs
    return n, cost, hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("binary")
    ap.add_argument("--func", default="")
    ap.add_argument("--measured", default=None)
    ap.add_argument("--dump", action="store_true", help="print the FP instructions of the loop")
    a = ap.parse_args()
    meas = {}
    if a.measured:
        for line in open(a.measured):
            m = re.match(r"(scalar|packed|mul)\s+N=(\d+) A=(\d+) B=(\d+).*ipc ([0-9.]+)", line)
            if m:
                kind = {"scalar": "bank_scalar", "packed": "bank_packed", "mul": "bank_mul"}[m.group(1)]
                meas[f"{kind}<{m.group(2)}, {m.group(3)}, {m.group(4)}>"] = float(m.group(5))
    print(f"{'kernel':58s} {'FP inst':>7s} {'port2':>7s} {'bank2':>7s} {'bank4':>7s} {'measured':>9s}   (warp-inst/clk/SMSP over the loop's FP instructions)")
    for name, ins in functions(a.binary):
        dn = demangle(name)
        if a.func and a.func not in dn:
            continue
        body = hot_loop(ins)
        if not body:
            continue
        n, cost, hist = model(body)
        if n == 0:
            continue
        short = re.sub(r"^void ", "", dn).split("(")[0]
        mv = next((v for k, v in meas.items() if k in dn), None)
        print(f"{short[:58]:58s} {n:7d} {n / cost['port2']:7.3f} {n / cost['bank2']:7.3f} {n / cost['bank4']:7.3f} "
              f"{(f'{mv:9.3f}' if mv is not None else '        -')}")
        if a.dump:
            print("   loop: %d instructions; (opcode, RF reads) histogram: %s" % (len(body), dict(sorted(hist.items()))))
            print(f"   predicted cycles per loop iteration per warp: port2 {cost['port2']}  bank2 {cost['bank2']}  bank4 {cost['bank4']}"
                  f"  (non-FP instructions: {len(body) - n})")


if __name__ == "__main__":
    main()
