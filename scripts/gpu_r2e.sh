#!/bin/bash
# round-2 GPU session E: fp16 conv + split GEMM epilogue + deterministic reductions: full suite, bench, launch list
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2e_suite.log; tail -12 gpurun_out/r2e_suite.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:14]: print(k, v)
PY
for spec in "tc_gemm_kernel:70:3" "conv_fwd_mma16_kernel:70:1"; do
  pat=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 420 ncu --set full --import-source on --clock-control none -k regex:$pat --launch-skip $skip --launch-count $cnt \
    -o gpurun_out/r2e_ncu_$pat -f python bench.py --steps 1 --warmup 1 --no-cpu --no-env-roofline > gpurun_out/r2e_ncu_$pat.log 2>&1
  ls -la gpurun_out/r2e_ncu_$pat.ncu-rep 2>/dev/null
done
