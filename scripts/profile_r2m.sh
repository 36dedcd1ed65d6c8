#!/bin/bash
# round-2 final ncu evidence on ONE B200: (1) the launch list of two steady-state updates of the default bench command
# (per-launch gpu__time_duration, cold-cache and serialised: shares, not absolutes), (2) one `--set full` capture of
# every kernel family of the training step and of the rollout variants.  Only the conv / GEMM reports are kept as
# .ncu-rep (64 MiB pull limit); the others are summarised on the box.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-env-roofline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 6000 --csv \
    --log-file gpurun_out/r2m_launches.csv $BENCH > gpurun_out/r2m_launches_bench.log 2>&1
for spec in "conv_bwd_mma16_kernel:70:1" "conv_fwd_mma16_kernel:140:1" "tc_gemm_kernel:140:3" "row_bwd_kernel:70:1" "radam_kernel:70:1" "rollout_act_step_kernel:5:1"; do
  pat=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 420 ncu --set full --import-source on --clock-control none -k regex:$pat --launch-skip $skip --launch-count $cnt \
    -o gpurun_out/r2m_ncu_$pat -f $BENCH > gpurun_out/r2m_ncu_$pat.log 2>&1
  ls -la gpurun_out/r2m_ncu_$pat.ncu-rep 2>/dev/null
done
# rollout variants: the first launches of the families are the rollout forwards (inference conv, Q-head GEMM)
timeout 420 ncu --set full --import-source on --clock-control none -k regex:conv_fwd_mma16_kernel --launch-skip 5 --launch-count 1 \
    -o gpurun_out/r2m_ncu_rollout_conv -f $BENCH > gpurun_out/r2m_ncu_rollout_conv.log 2>&1
timeout 420 ncu --set full --import-source on --clock-control none -k regex:tc_gemm_kernel --launch-skip 5 --launch-count 1 \
    -o gpurun_out/r2m_ncu_rollout_gemm -f $BENCH > gpurun_out/r2m_ncu_rollout_gemm.log 2>&1
python scripts/ncu_rep_summary.py 524288 gpurun_out/r2m_ncu_*.ncu-rep > gpurun_out/r2m_ncu_summary.md 2> gpurun_out/r2m_ncu_summary.err
python scripts/make_traffic_json.py gpurun_out/r2m_traffic.json gpurun_out/r2m_ncu_*.ncu-rep > /dev/null 2>> gpurun_out/r2m_ncu_summary.err
rm -f gpurun_out/r2m_ncu_rollout_*.ncu-rep gpurun_out/r2m_ncu_radam_kernel.ncu-rep gpurun_out/r2m_ncu_rollout_act_step_kernel.ncu-rep gpurun_out/r2m_ncu_row_bwd_kernel.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | grep r2m
