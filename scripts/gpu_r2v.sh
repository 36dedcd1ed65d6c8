#!/bin/bash
# round-2 final 1-GPU session V: full GPU suite, default bench (the driver's command), reference arm, Acrobot / MinAtar
# lines, and a refreshed ncu launch list + conv captures of the final kernels
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python -m pytest tests -q -m gpu --deselect tests/test_gpu_multi.py 2>&1 | tail -12 > gpurun_out/r2v_pytest_gpu.log; tail -3 gpurun_out/r2v_pytest_gpu.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; tail -2 gpurun_out/r2v_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference_arm.json 2> gpurun_out/r2v_bench_reference_arm.err; tail -2 gpurun_out/r2v_bench_reference_arm.err; cut -c1-300 gpurun_out/r2v_bench_reference_arm.json
python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/r2v_bench_8steps.json 2> gpurun_out/r2v_bench_8steps.err
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2v_bench_acrobot.json 2> gpurun_out/r2v_bench_acrobot.err
python bench.py --config minatar5 --steps 10 --warmup 3 > gpurun_out/r2v_bench_minatar5.json 2> gpurun_out/r2v_bench_minatar5.err
python bench.py --with-eval --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2v_bench_with_eval.json 2> gpurun_out/r2v_bench_with_eval.err
python - <<'PY'
import json
for f in ("r2v_bench","r2v_bench_8steps","r2v_bench_with_eval","r2v_bench_acrobot"):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e6,2), "M", round(d["ms_per_step"],1), "ms  e2e", round(d["e2e"]["value"]/1e6,2), d["clocks"]["sm_mhz"], (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
for l in open('gpurun_out/r2v_bench_minatar5.json'):
    d=json.loads(l); print(d["metric"][:40], round(d["value"]/1e6,2))
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k,v)
PY
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu --no-env-roofline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 6000 --csv --log-file gpurun_out/r2v_launches.csv $BENCH > gpurun_out/r2v_launches_bench.log 2>&1
for spec in "conv_bwd_mma16_kernel:70:1" "conv_fwd_mma16_kernel:140:1"; do
  pat=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 420 ncu --set full --import-source on --clock-control none -k regex:$pat --launch-skip $skip --launch-count $cnt \
    -o gpurun_out/r2v_ncu_$pat -f $BENCH > gpurun_out/r2v_ncu_$pat.log 2>&1
done
python scripts/ncu_rep_summary.py 524288 gpurun_out/r2v_ncu_*.ncu-rep > gpurun_out/r2v_ncu_summary.md 2> gpurun_out/r2v_ncu_summary.err
ls -la gpurun_out | grep r2v | head -30
