#!/usr/bin/env python
"""Markdown summary of `ncu --set full` reports (one row per captured launch) + SASS opcode histogram:
    python scripts/ncu_rep_summary.py <units per launch> a.ncu-rep [b.ncu-rep ...] > profiles/x.md"""
import collections
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "HMMA pipe %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor cycles %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst")]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def opcodes(path, units):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, cur = {}, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = res.setdefault(r[1][:70], collections.Counter())
            hdr = None
            continue
        if r and r[0] == "Address":
            hdr = r
            src, ie = hdr.index("Source"), hdr.index("Instructions Executed")
            continue
        if cur is None or hdr is None or len(r) <= ie:
            continue
        try:
            n = int(r[ie])
        except ValueError:
            continue
        toks = r[src].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        cur[op.split(".")[0]] += n
    return res


def main():
    units = float(sys.argv[1])
    for path in sys.argv[2:]:
        rows, un = raw(path)
        print(f"## {path.split('/')[-1]}\n")
        print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
        print("|---|" + "---|" * len(KEYS))
        for d in rows:
            cells = []
            for k, _ in KEYS:
                v = d.get(k, "")
                try:
                    v = f"{float(v):.4g} {un.get(k, '')}".strip()
                except ValueError:
                    pass
                cells.append(v)
            print(f"| `{d['Kernel Name'][:60]}` | " + " | ".join(cells) + " |")
        st = {}
        for d in rows[:1]:
            st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
                  for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and v}
        if st:
            print("\nwarp stalls per issued instruction (first launch): " +
                  ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
        for name, c in opcodes(path, units).items():
            tot = sum(c.values())
            print(f"\nSASS opcodes of `{name}`: {tot / units:.0f} warp-instructions per unit ({units:.0f} units/launch): " +
                  ", ".join(f"{k} {v / units:.0f}" for k, v in c.most_common(16)))
        print()


if __name__ == "__main__":
    main()
