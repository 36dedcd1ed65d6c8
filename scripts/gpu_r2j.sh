#!/bin/bash
# round-2 GPU session J (1 GPU): thin first-layer weight gradient + parallel finalize kernels; 1-seed x 65,536-env leg
mkdir -p gpurun_out
python -m pytest tests/test_gpu_net.py tests/test_gpu_norm.py tests/test_gpu_train.py tests/test_gpu_parity_r2.py tests/test_gpu_rnn.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r2j_tests.log; tail -4 gpurun_out/r2j_tests.log
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2j_bench_acrobot.json 2> gpurun_out/r2j_bench_acrobot.err; tail -2 gpurun_out/r2j_bench_acrobot.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench_acrobot.json').read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k,v)
PY
python bench.py --gpus 1 --seeds 1 --envs 65536 --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2j_bench_1seed_65536env_1gpu.json 2> gpurun_out/r2j_bench_1seed_65536env_1gpu.err; tail -2 gpurun_out/r2j_bench_1seed_65536env_1gpu.err; cut -c1-330 gpurun_out/r2j_bench_1seed_65536env_1gpu.json
