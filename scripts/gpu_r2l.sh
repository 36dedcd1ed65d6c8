#!/bin/bash
# round-2 GPU session L (2 GPUs): env-sharded run against a single-rank run, env-sharded bench at 1 seed x 65,536 envs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2l_multi.log; tail -4 gpurun_out/r2l_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --seeds 1 --envs 65536 --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2l_bench_1seed_65536env_2gpu_envsharded.json 2> gpurun_out/r2l_bench_2gpu.err; tail -3 gpurun_out/r2l_bench_2gpu.err; cut -c1-330 gpurun_out/r2l_bench_1seed_65536env_2gpu_envsharded.json
