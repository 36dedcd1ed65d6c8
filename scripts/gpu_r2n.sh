#!/bin/bash
# round-2 GPU session N (1 GPU): two-deep gather prefetch in the conv forward, 32-slice finalize kernels for few-seed runs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_net.py tests/test_gpu_train.py tests/test_gpu_parity_r2.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r2n_tests.log; tail -4 gpurun_out/r2n_tests.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], d["clocks"], d["roofline"]["frac"])
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k,v)
PY
python bench.py --gpus 1 --seeds 1 --envs 65536 --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2n_bench_1seed_65536env_1gpu.json 2> gpurun_out/r2n_bench_1seed.err; tail -2 gpurun_out/r2n_bench_1seed.err
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2n_bench_acrobot.json 2> gpurun_out/r2n_bench_acrobot.err; tail -2 gpurun_out/r2n_bench_acrobot.err
python - <<'PY'
import json
for f in ("r2n_bench_1seed_65536env_1gpu","r2n_bench_acrobot"):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"])
    for k,v in list(d["kernel_breakdown"].items())[:8]: print("  ",k,v)
PY
