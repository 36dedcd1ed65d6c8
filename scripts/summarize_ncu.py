#!/usr/bin/env python
"""Summarise ncu outputs (no GPU needed): per-kernel share of the launch list and the
key metrics of each --set full capture.   python scripts/summarize_ncu.py gpurun_out r1 > profiles/r1_ncu_summary.md"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

d, tag = sys.argv[1], sys.argv[2]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed"]
print(f"# ncu summary ({tag})\n")
lf = os.path.join(d, f"{tag}_launches.csv")
if os.path.exists(lf):
    rows = [l for l in open(lf) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    tot = defaultdict(float); cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", r["Kernel Name"]).replace("pqn::", "")
        name = re.sub(r"\(.*", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        tot[name] += v; cnt[name] += 1
    allt = sum(tot.values())
    print("## launch list of one timed step (`ncu --metrics gpu__time_duration.sum`, cold-cache serialised: compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"| {k} | {cnt[k]} | {tot[k]:.0f} | {tot[k] / allt:.3f} |")
    print(f"\ntotal {allt / 1e3:.1f} ms over {sum(cnt.values())} launches\n")
for rep in sorted(glob.glob(os.path.join(d, f"{tag}_*.ncu-rep"))):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"## {os.path.basename(rep)} — {m.get('Kernel Name', ('?',))[0][:90]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in m:
            print(f"| {k} | {m[k][0]} | {m[k][1]} |")
    print()
