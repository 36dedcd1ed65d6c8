#!/bin/bash
# same-box A/B of two library builds: scripts/ab/libpqn_old.so (A) against the in-tree build (B), interleaved A B A B
mkdir -p gpurun_out
cp purejaxql_b200/libpqn_b200.so /tmp/new.so
run() {  # $1 = tag
  python bench.py --steps 6 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - "$1" <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/ab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
kb=d["kernel_breakdown"]
print(sys.argv[1], round(d["value"]/1e6,2), "M", round(d["ms_per_step"],1), "ms", d["clocks"]["sm_mhz"], {k: kb[k]["ms_per_update"] for k in ("conv_fwd","conv_bwd","tc_dgrad","conv_fwd_infer","tc_dense_fwd")})
PY
}
cp scripts/ab/libpqn_old.so purejaxql_b200/libpqn_b200.so; run A1
cp /tmp/new.so purejaxql_b200/libpqn_b200.so; run B1
cp scripts/ab/libpqn_old.so purejaxql_b200/libpqn_b200.so; run A2
cp /tmp/new.so purejaxql_b200/libpqn_b200.so; run B2
