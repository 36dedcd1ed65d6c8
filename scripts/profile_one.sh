#!/bin/bash
# usage: profile_one.sh <tag> <kernel-regex> <skip> <count>   (run under gpurun; exports CSV pages, drops the .ncu-rep)
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu"
stem=gpurun_out/$1
ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o $stem -f $BENCH > $stem.log 2>&1
ncu -i $stem.ncu-rep --page raw --csv > $stem.raw.csv 2>/dev/null
ncu -i $stem.ncu-rep --page source --csv 2>/dev/null | gzip > $stem.source.csv.gz
rm -f $stem.ncu-rep
