#!/bin/bash
# one more run of the driver's default command on whatever box comes up (box-to-box clock spread)
mkdir -p gpurun_out
python bench.py > gpurun_out/r2z2_bench.json 2> gpurun_out/r2z2_bench.err; tail -2 gpurun_out/r2z2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z2_bench.json').read().strip().splitlines()[-1])
print(round(d["value"]/1e6,2), "M", round(d["ms_per_step"],1), "ms  e2e", round(d["e2e"]["value"]/1e6,2), d["clocks"], d["roofline"]["kernel"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
PY
