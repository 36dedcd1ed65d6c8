#!/bin/bash
# round-2 final 2-GPU session W: env-sharded vs single-rank parity test, seed-sharded headline on 2 GPUs, env-sharded
# 1 seed x 65,536 envs on 2 GPUs (+ the same workload on one GPU for the ratio)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -6 > gpurun_out/r2w_pytest_gpu_multi.log; tail -3 gpurun_out/r2w_pytest_gpu_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2w_bench_2gpu.json 2> gpurun_out/r2w_bench_2gpu.err; tail -2 gpurun_out/r2w_bench_2gpu.err; cut -c1-260 gpurun_out/r2w_bench_2gpu.json
CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --seeds 1 --envs 65536 --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2w_bench_1seed_65536env_1gpu.json 2> gpurun_out/r2w_bench_1seed_1gpu.err; cut -c1-260 gpurun_out/r2w_bench_1seed_65536env_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --seeds 1 --envs 65536 --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2w_bench_1seed_65536env_2gpu_envsharded.json 2> gpurun_out/r2w_bench_1seed_2gpu.err; tail -2 gpurun_out/r2w_bench_1seed_2gpu.err; cut -c1-260 gpurun_out/r2w_bench_1seed_65536env_2gpu_envsharded.json
