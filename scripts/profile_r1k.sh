#!/bin/bash
# Round-1 (second half) ncu evidence, run under gpurun on ONE B200: launch list + full captures per kernel family.
# gpurun copies back at most 64 MiB, so each capture is exported to CSV on the box and only two .ncu-rep are kept.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 1000 --csv \
    --log-file gpurun_out/r1k_launches.csv $BENCH > gpurun_out/r1k_launches_bench.log 2>&1
export_rep() {  # $1 = report stem
  ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
  ncu -i $1.ncu-rep --page details --csv > $1.details.csv 2>/dev/null
  ncu -i $1.ncu-rep --page source --csv 2>/dev/null | gzip > $1.source.csv.gz
}
# training-phase launches (skip the rollout launches of the family first)
for k in tc_gemm_kernel conv_bwd_mma_kernel conv_fwd_mma_kernel row_bwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 3 \
      -o gpurun_out/r1k_$k -f $BENCH > gpurun_out/r1k_${k}.log 2>&1
  export_rep gpurun_out/r1k_$k
done
# rollout-phase variants (Q-head epilogue GEMM, inference conv)
for k in tc_gemm_kernel conv_fwd_mma_kernel rollout_act_step_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 \
      -o gpurun_out/r1k_rollout_$k -f $BENCH > gpurun_out/r1k_rollout_${k}.log 2>&1
  export_rep gpurun_out/r1k_rollout_$k
done
rm -f gpurun_out/r1k_conv_fwd_mma_kernel.ncu-rep gpurun_out/r1k_row_bwd_kernel.ncu-rep gpurun_out/r1k_rollout_*.ncu-rep
du -sh gpurun_out
ls -la gpurun_out | grep r1k
