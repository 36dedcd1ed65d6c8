#!/bin/bash
# Round-1 final ncu evidence (run under gpurun on ONE B200).
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 1000 --csv \
    --log-file gpurun_out/r1h_launches.csv $BENCH > gpurun_out/r1h_launches_bench.log 2>&1
# full captures: 3 consecutive launches of each kernel family from the timed step
for k in tc_gemm_kernel conv_bwd_mma_kernel conv_fwd_mma_kernel row_bwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 3 \
      -o gpurun_out/r1h_$k -f $BENCH > gpurun_out/r1h_${k}.log 2>&1
done
ls -la gpurun_out | grep r1h
