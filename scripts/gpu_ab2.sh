#!/bin/bash
# same-box sensitivity of the training conv forward: in-tree build (B) vs variants without xhat stores (V1), without
# fp16 plane stores (V2), without ReLU-bit gathering (V3).  The variants compute wrong results: timing only.
mkdir -p gpurun_out
cp purejaxql_b200/libpqn_b200.so /tmp/new.so
run() {
  python bench.py --steps 4 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/ab2_$1.json 2> gpurun_out/ab2_$1.err
  python - "$1" <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/ab2_{sys.argv[1]}.json').read().strip().splitlines()[-1])
kb=d["kernel_breakdown"]
print(sys.argv[1], round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], {k: kb[k]["ms_per_update"] for k in ("conv_fwd","conv_bwd","tc_dgrad","conv_fwd_infer")})
PY
}
run B1
for v in v1 v2 v3; do cp scripts/ab/libpqn_$v.so purejaxql_b200/libpqn_b200.so; run $v; done
cp /tmp/new.so purejaxql_b200/libpqn_b200.so; run B2
