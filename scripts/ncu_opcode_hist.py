#!/usr/bin/env python
"""Opcode histogram (warp instructions executed per sample) of one kernel from an `ncu --import-source on` report:
    ncu -i X.ncu-rep --page source --csv | python scripts/ncu_opcode_hist.py <units per launch>"""
import collections
import csv
import sys

units = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
src, ie, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
c, st, tot, tots = collections.Counter(), collections.Counter(), 0, 0
for r in rows[hi + 1:]:
    try:
        n, s = int(r[ie]), int(r[ss])
    except (ValueError, IndexError):
        continue
    toks = r[src].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    c[op] += n; st[op] += s; tot += n; tots += s
print(f"total warp-instructions {tot}  per unit {tot / units:.1f}   stall samples {tots}")
for k, v in c.most_common(28):
    print(f"{k:10s} {v / units:8.1f} per unit  {100 * v / tot:5.1f} % of instr   {100 * st[k] / max(tots, 1):5.1f} % of samples")
