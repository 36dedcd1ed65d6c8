#!/bin/bash
# round-2 GPU session Q (1 GPU): A-stationary dgrad (two-stage ring, A tiles loaded once per m-tile group)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2q_tc.log; tail -4 gpurun_out/r2q_tc.log
timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu -x -k "loss_grad" 2>&1 | tail -8 > gpurun_out/r2q_net.log; tail -3 gpurun_out/r2q_net.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -2 gpurun_out/r2q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d["roofline"]["kernel"], d["roofline"]["frac"], "td", d["td_loss_last"])
for k,v in list(d["kernel_breakdown"].items())[:10]: print(k,v)
PY
