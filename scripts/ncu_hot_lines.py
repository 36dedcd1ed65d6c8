#!/usr/bin/env python
"""Top SASS lines by stall samples of an `ncu --set full --import-source on` report:
    python scripts/ncu_hot_lines.py report.ncu-rep [kernel-substring] [N]"""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    # the export concatenates kernels: a "Kernel Name" row starts each block
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["data"].append(r)
    for b in blocks:
        if want not in b["name"]:
            continue
        ix = {h: i for i, h in enumerate(b["hdr"])}

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        tot = sum(f(r, "# Samples") for r in b["data"]) or 1.0
        print("##", b["name"][:100], "samples", int(tot))
        keys = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
        agg = sorted(((sum(f(r, k) for r in b["data"]), k) for k in keys), reverse=True)
        print("  " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for v, k in agg[:8]))
        for r in sorted(b["data"], key=lambda r: -f(r, "# Samples"))[:topn]:
            best = max(keys, key=lambda k: f(r, k))
            print(f"  {100 * f(r, '# Samples') / tot:5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:80]:80s} {best[6:]}")


if __name__ == "__main__":
    main()
