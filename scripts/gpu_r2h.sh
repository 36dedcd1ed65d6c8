#!/bin/bash
# round-2 GPU session H (2 GPUs): env-sharded mode against a single-rank run, seed-sharded and env-sharded bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests/test_gpu_net.py tests/test_gpu_norm.py tests/test_gpu_tc.py -q -m gpu 2>&1 | tail -12 > gpurun_out/r2h_tests.log; tail -5 gpurun_out/r2h_tests.log
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2h_bench_acrobot.json 2> gpurun_out/r2h_bench_acrobot.err; tail -2 gpurun_out/r2h_bench_acrobot.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_acrobot.json').read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:8]: print(k,v)
PY
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2h_multi.log; tail -8 gpurun_out/r2h_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_2gpu.json 2> gpurun_out/r2h_bench_2gpu.err; tail -3 gpurun_out/r2h_bench_2gpu.err; cut -c1-400 gpurun_out/r2h_bench_2gpu.json
python bench.py --gpus 1 --seeds 1 --steps 20 --warmup 5 --no-cpu --no-env-roofline > gpurun_out/r2h_bench_1seed_1gpu.json 2> gpurun_out/r2h_bench_1seed_1gpu.err; cut -c1-330 gpurun_out/r2h_bench_1seed_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --seeds 1 --steps 20 --warmup 5 --no-env-roofline > gpurun_out/r2h_bench_1seed_2gpu_envsharded.json 2> gpurun_out/r2h_bench_1seed_2gpu.err; tail -3 gpurun_out/r2h_bench_1seed_2gpu.err; cut -c1-330 gpurun_out/r2h_bench_1seed_2gpu_envsharded.json
