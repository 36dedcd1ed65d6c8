#!/bin/bash
# round-2 GPU session M (1 GPU): ncu launch list + full captures of the final kernels, sanitizer passes, learning curves
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python scripts/learn_check.py 3e6 > gpurun_out/r2m_learning_breakout.json 2> gpurun_out/r2m_learning_breakout.err; cut -c1-500 gpurun_out/r2m_learning_breakout.json
python scripts/learn_all.py > gpurun_out/r2m_learning_all_envs.jsonl 2> gpurun_out/r2m_learning_all.err; cut -c1-300 gpurun_out/r2m_learning_all_envs.jsonl
bash scripts/sanitize_r2.sh 2>&1 | tail -24
bash scripts/profile_r2m.sh 2>&1 | tail -25
