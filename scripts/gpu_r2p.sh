#!/bin/bash
# round-2 GPU session P (1 GPU): conv backward with the weight gradient on fp16 mma (exponent-coded pair words)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py -q -m gpu -x -k "loss_grad" 2>&1 | tail -15 > gpurun_out/r2p_net.log; tail -6 gpurun_out/r2p_net.log
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity_r2.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2p_tests.log; tail -4 gpurun_out/r2p_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -2 gpurun_out/r2p_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d["roofline"]["frac"])
for k,v in list(d["kernel_breakdown"].items())[:10]: print(k,v)
PY
