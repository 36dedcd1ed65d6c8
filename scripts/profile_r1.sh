#!/bin/bash
# Round-1 ncu evidence (run under gpurun on ONE B200).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu"
# 1) every launch of the timed step with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 780 -c 760 --csv \
    --log-file gpurun_out/r1_launches.csv $BENCH > gpurun_out/r1_launches_bench.log 2>&1
# 2) full-set captures of the top kernels (1 launch each, from the timed step)
for k in dense_fwd_kernel dgrad_kernel wgrad_kernel conv_bwd_kernel conv_fwd_kernel rollout_act_step_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 \
      -o gpurun_out/r1_$k -f $BENCH > gpurun_out/r1_${k}.log 2>&1
done
python scripts/bench_env_step.py > gpurun_out/r1_env_step_bench.json 2> gpurun_out/r1_env_step_bench.err
ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 6 -c 1 \
    -o gpurun_out/r1_env_step_kernel -f python scripts/bench_env_step.py > gpurun_out/r1_env_step_ncu.log 2>&1
ls -la gpurun_out
