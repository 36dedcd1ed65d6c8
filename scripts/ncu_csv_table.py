#!/usr/bin/env python
"""Per-launch markdown table + per-kernel DRAM traffic JSON from the raw-page CSVs that scripts/profile_*.sh
export on the GPU box (no GPU needed).  usage: ncu_csv_table.py <dir> <tag>   -> stdout (markdown),
profiles/<tag>_traffic.json"""
import csv, glob, json, os, sys
d, tag = sys.argv[1], sys.argv[2]
KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__inst_executed.sum", "warp_inst"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
print(f"# ncu --set full --clock-control none captures ({tag}): one row per captured launch\n")
print("Command: `python bench.py --steps 1 --warmup 1 --no-cpu` (S=128 seeds x 4096 envs), see scripts/profile_%s.sh\n" % tag)
print("| capture | kernel | " + " | ".join(k for _, k in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
traffic = {}
for f in sorted(glob.glob(os.path.join(d, f"{tag}_*.raw.csv"))):
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        short = m.get("Kernel Name", "?").replace("pqn::", "").replace("tc::", "").split("(")[0][-44:]
        cells = []
        for k, _ in KEYS:
            v = m.get(k, "")
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u.get(k, '')}".strip())
        print(f"| {os.path.basename(f)[len(tag) + 1:-8]} | {short} | " + " | ".join(cells) + " |")
        try:
            rd = float(m["dram__bytes_read.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_read.sum"], 1)
            wr = float(m["dram__bytes_write.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_write.sum"], 1)
            traffic.setdefault(short, []).append(rd + wr)
        except Exception:
            pass
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(os.path.join("profiles", f"{tag}_traffic.json"), "w"),
          indent=1)
