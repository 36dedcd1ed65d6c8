#!/bin/bash
# compute-sanitizer passes over small GPU tests (run under gpurun on one B200); logs -> gpurun_out/
mkdir -p gpurun_out
K="test_cnn_loss_grad and inline_lo and mma_conv and 64"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_net.py -x -q -m gpu -k "$K" > gpurun_out/memcheck_net.log 2>&1
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_net.py -x -q -m gpu -k "$K" > gpurun_out/racecheck_net.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -x -q -m gpu \
  -k "test_minatar_update_step_matches_oracle or test_gymnax_update_step_matches_oracle" > gpurun_out/memcheck_update.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_env.py -x -q -m gpu \
  -k "golden or eps_greedy or empty" > gpurun_out/memcheck_env.log 2>&1
tail -n 2 gpurun_out/memcheck_*.log gpurun_out/racecheck_*.log
