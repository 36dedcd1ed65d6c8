#!/usr/bin/env python
"""Per-launch table of every kernel instance inside the ncu --set full captures (no GPU needed)."""
import csv, glob, io, os, subprocess, sys, json
d, tag = sys.argv[1], sys.argv[2]
KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]
print(f"# ncu --set full captures ({tag}): one row per captured launch\n")
print("| file | kernel | " + " | ".join(k for _, k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
traffic = {}
for rep in sorted(glob.glob(os.path.join(d, f"{tag}_*.ncu-rep"))):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        name = m.get("Kernel Name", "?").replace("pqn::", "")
        short = name.split("(")[0][-42:]
        cells = []
        for k, _ in KEYS:
            v = m.get(k, "")
            cells.append(f"{v} {u.get(k, '')}".strip())
        print(f"| {os.path.basename(rep)} | {short} | " + " | ".join(cells) + " |")
        def tobytes(v, unit):
            try:
                x = float(v.replace(",", ""))
            except Exception:
                return None
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        rd, wr = tobytes(m.get("dram__bytes_read.sum", ""), u.get("dram__bytes_read.sum", "")), tobytes(m.get("dram__bytes_write.sum", ""), u.get("dram__bytes_write.sum", ""))
        if rd is not None and wr is not None:
            traffic.setdefault(short, []).append(rd + wr)
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(os.path.join("profiles", f"{tag}_traffic.json"), "w"), indent=1)
