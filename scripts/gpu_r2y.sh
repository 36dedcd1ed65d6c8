#!/bin/bash
# round-2 GPU session Y (1 GPU): conv forward with permuted output channels (16-byte xhat / 8-byte plane stores per thread)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_train.py tests/test_gpu_parity_r2.py tests/test_gpu_norm.py -q -m gpu 2>&1 | tail -6 > gpurun_out/r2y_tests.log; tail -4 gpurun_out/r2y_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -2 gpurun_out/r2y_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d["roofline"]["kernel"], d["roofline"]["frac"], d["td_loss_last"])
for k,v in list(d["kernel_breakdown"].items())[:10]: print(k,v)
PY
