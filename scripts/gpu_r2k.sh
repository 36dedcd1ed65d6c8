#!/bin/bash
# round-2 GPU session K (1 GPU): deterministic FFMA weight gradients, vectorised thin kernel, determinism of every variant
mkdir -p gpurun_out
python -m pytest tests/test_gpu_net.py tests/test_gpu_norm.py tests/test_gpu_train.py tests/test_gpu_rnn.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r2k_tests.log; tail -4 gpurun_out/r2k_tests.log
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2k_bench_acrobot.json 2> gpurun_out/r2k_bench_acrobot.err; tail -2 gpurun_out/r2k_bench_acrobot.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench_acrobot.json').read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k,v)
PY
