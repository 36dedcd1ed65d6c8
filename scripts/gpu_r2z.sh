#!/bin/bash
# round-2 closing 1-GPU session Z (final code): full GPU suite, the driver's default bench + reference arm, Acrobot and
# MinAtar lines, launch list of the default command, ncu of the two conv kernels
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests -q -m gpu --deselect tests/test_gpu_multi.py 2>&1 | tail -12 > gpurun_out/r2z_pytest_gpu.log; tail -3 gpurun_out/r2z_pytest_gpu.log
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -2 gpurun_out/r2z_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2z_bench_acrobot.json 2> gpurun_out/r2z_bench_acrobot.err
python bench.py --config minatar5 --steps 40 --warmup 3 > gpurun_out/r2z_bench_minatar5.json 2> gpurun_out/r2z_bench_minatar5.err
python - <<'PY'
import json
for f in ("r2z_bench","r2z_bench_acrobot"):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2), "M", round(d["ms_per_step"],1), "ms  e2e", round(d["e2e"]["value"]/1e6,2), d["clocks"]["sm_mhz"], (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"), d["e2e"].get("wall_split_rank0"))
    except Exception as e: print(f, "ERR", e)
for l in open('gpurun_out/r2z_bench_minatar5.json'):
    d=json.loads(l); print(d["metric"][:40], round(d["value"]/1e6,2))
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k,v)
PY
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-env-roofline"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3600 --csv --log-file gpurun_out/r2z_launches.csv $BENCH > gpurun_out/r2z_launches_bench.log 2>&1
for spec in "conv_bwd_mma16_kernel:70:1" "conv_fwd_mma16_kernel:140:1"; do
  pat=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$pat --launch-skip $skip --launch-count $cnt \
    -o gpurun_out/r2z_ncu_$pat -f $BENCH > gpurun_out/r2z_ncu_$pat.log 2>&1
done
python scripts/ncu_rep_summary.py 524288 gpurun_out/r2z_ncu_*.ncu-rep > gpurun_out/r2z_ncu_summary.md 2> gpurun_out/r2z_ncu_summary.err
python scripts/make_traffic_json.py gpurun_out/r2z_traffic.json gpurun_out/r2z_ncu_*.ncu-rep > /dev/null 2>&1
ls gpurun_out | grep r2z | head -30
