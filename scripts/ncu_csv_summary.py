#!/usr/bin/env python
"""Summarise the CSV pages exported by scripts/profile_one.sh / profile_r1i.sh (no GPU needed).
usage: ncu_csv_summary.py <stem> [<stem> ...]   (stem = gpurun_out/<tag>)"""
import collections, csv, gzip, sys
KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
for stem in sys.argv[1:]:
    rows = list(csv.reader(open(stem + ".raw.csv")))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print(stem.split("/")[-1], m.get("Kernel Name", "?").replace("pqn::", "")[:50])
        print("   ", " ".join(f"{k}={m.get(K, '')}{u.get(K, '')[:5]}" for K, k in KEYS))
    try:
        rows = list(csv.reader(gzip.open(stem + ".source.csv.gz", "rt")))
    except Exception:
        continue
    # one block per kernel instance: header row starts with "Address"
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Address":
            cur = (r, []); blocks.append(cur)
        elif cur is not None and len(r) == len(cur[0]):
            cur[1].append(dict(zip(cur[0], r)))
    for hdr, rs in blocks[:1]:
        tot = sum(int(r["Instructions Executed"]) for r in rs); samp = sum(int(r["# Samples"]) for r in rs)
        print(f"    SASS lines {len(rs)}  warp-instructions {tot}  samples {samp}")
        h = collections.Counter(); st = collections.Counter()
        for r in rs:
            toks = r["Source"].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            h[op] += int(r["Instructions Executed"]); st[op] += int(r["# Samples"])
        print("    " + "  ".join(f"{k}:{v * 100 // tot}%i/{st[k] * 100 // max(samp, 1)}%s" for k, v in h.most_common(14)))
        stalls = collections.Counter()
        for r in rs:
            for k, v in r.items():
                if k.startswith("stall_") and "Not Issued" not in k and v.isdigit():
                    stalls[k] += int(v)
        print("    stalls: " + "  ".join(f"{k[6:]}:{v * 100 // max(samp, 1)}%" for k, v in stalls.most_common(8)))
