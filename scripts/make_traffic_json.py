#!/usr/bin/env python
"""profiles/r2_traffic.json (read by bench.py for `roofline.traffic`) from `ncu --set full` reports:
    python scripts/make_traffic_json.py out.json a.ncu-rep [b.ncu-rep ...]
dram__bytes_read.sum + dram__bytes_write.sum of the first captured launch of every training-step kernel family."""
import csv
import json
import subprocess
import sys

NAMES = {"tc_gemm_kernel<1, 1, 0, 1>": "tc_wgrad", "tc_gemm_kernel<0, 0, 4, 1>": "tc_dgrad",
         "tc_gemm_kernel<0, 1, 1, 1>": "tc_dense_fwd", "tc_gemm_kernel<0, 1, 2, 1>": "tc_dense_fwd_head",
         "conv_fwd_mma16_kernel<4, 1, 1>": "conv_fwd", "conv_fwd_mma16_kernel<4, 0, 1>": "conv_fwd_infer",
         "conv_bwd_mma16_kernel<4>": "conv_bwd", "conv_bwd_mma_kernel<4>": "conv_bwd_tf32", "row_bwd_kernel<128, 1>": "row_bwd",
         "radam_kernel": "radam", "rollout_act_step_kernel": "rollout_act_step"}
MULT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    res, src = {}, []
    for path in reps:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], dict(zip(rows[0], rows[1]))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            for pat, name in NAMES.items():
                if pat in d["Kernel Name"] and name not in res:
                    res[name] = sum(float(d[k]) * MULT[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        src.append(path.split("/")[-1])
    res["_source"] = ("ncu --set full --clock-control none, one steady-state launch each at S=128, E=4096 "
                      "(bench.py --steps 1 --warmup 1 --no-cpu --no-env-roofline): " + ", ".join(src) +
                      "; dram__bytes_read.sum + dram__bytes_write.sum")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
