#!/bin/bash
# round-2 compute-sanitizer passes over small GPU tests of the default paths (run under gpurun on one B200)
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 9"
$CS --tool memcheck python -m pytest tests/test_gpu_net.py -x -q -m gpu \
  -k "(test_cnn_loss_grad and f16split+f16_mma_conv and 64-200) or (test_mlp_loss_grad and hidden_layer_tcgen05 and 515) or (test_cnn_forward_matches and f16split+f16_mma_conv-130)" > gpurun_out/r2m_memcheck_net.log 2>&1
$CS --tool racecheck python -m pytest tests/test_gpu_net.py -x -q -m gpu \
  -k "(test_cnn_loss_grad and f16split+f16_mma_conv and 64-200) or (test_mlp_loss_grad and hidden_layer_tcgen05 and 4-128-1-2-100)" > gpurun_out/r2m_racecheck_net.log 2>&1
$CS --tool memcheck python -m pytest tests/test_gpu_norm.py -x -q -m gpu \
  -k "loss_grad and ((batch_norm-True) or (none-False))" > gpurun_out/r2m_memcheck_norm.log 2>&1
$CS --tool memcheck python -m pytest tests/test_gpu_rnn.py -x -q -m gpu -k "loss_grad or step_matches" > gpurun_out/r2m_memcheck_rnn.log 2>&1
$CS --tool memcheck python -m pytest tests/test_gpu_train.py -x -q -m gpu \
  -k "test_minatar_update_step_matches_oracle or test_gymnax_update_step_matches_oracle" > gpurun_out/r2m_memcheck_update.log 2>&1
$CS --tool memcheck python -m pytest tests/test_gpu_env.py -x -q -m gpu -k "golden or eps_greedy or empty" > gpurun_out/r2m_memcheck_env.log 2>&1
tail -n 2 gpurun_out/r2m_memcheck_*.log gpurun_out/r2m_racecheck_*.log
