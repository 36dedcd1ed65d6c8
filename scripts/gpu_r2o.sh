#!/bin/bash
# round-2 GPU session O (1 GPU): pqn_permutation (bucket + rank sort) instead of torch.sort: parity tests, suite, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_perm.py tests/test_gpu_env.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2o_perm.log; tail -4 gpurun_out/r2o_perm.log
python -m pytest tests/test_gpu_train.py tests/test_gpu_parity_r2.py tests/test_gpu_rnn.py tests/test_gpu_norm.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2o_tests.log; tail -4 gpurun_out/r2o_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; tail -2 gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d["roofline"]["frac"])
tot=sum(v["ms_per_update"] for v in d["kernel_breakdown"].values()); print("sum kernels", tot, "gap", d["ms_per_step"]-tot)
for k,v in list(d["kernel_breakdown"].items())[:14]: print(k,v)
PY
