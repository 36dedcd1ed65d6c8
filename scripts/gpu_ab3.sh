#!/bin/bash
# same-box A/B of conv forward launch shapes: in-tree 4 warps x 5 CTAs/SM (96 regs) vs v4 = 4 x 4 (119 regs),
# v6 = 8 warps x 2 CTAs (119 regs), v7 = 4 x 4 with the two m-block pairs unrolled (128 regs)
mkdir -p gpurun_out
cp purejaxql_b200/libpqn_b200.so /tmp/new.so
run() {
  python bench.py --steps 4 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/ab3_$1.json 2> gpurun_out/ab3_$1.err
  python - "$1" <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/ab3_{sys.argv[1]}.json').read().strip().splitlines()[-1])
kb=d["kernel_breakdown"]
print(sys.argv[1], round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], {k: kb[k]["ms_per_update"] for k in ("conv_fwd","conv_fwd_infer","conv_bwd")})
PY
}
run B1
for v in v4 v6 v7; do cp scripts/ab/libpqn_$v.so purejaxql_b200/libpqn_b200.so; run $v; done
cp /tmp/new.so purejaxql_b200/libpqn_b200.so; run B2
python bench.py --no-cpu --no-env-roofline > gpurun_out/ab3_default_k3.json 2> gpurun_out/ab3_default_k3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ab3_default_k3.json').read().strip().splitlines()[-1])
print("default K=3:", round(d["value"]/1e6,2), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]/1e6,2), d["e2e"]["wall_split_rank0"])
PY
