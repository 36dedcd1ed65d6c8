#!/bin/bash
# round-2 GPU session C: full GPU suite (norm variants included), ncu --set full of the hot kernels, bench
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_norm.py 2>&1 | tail -15 > gpurun_out/r2c_suite.log; tail -6 gpurun_out/r2c_suite.log
python -m pytest tests/test_gpu_norm.py -q -m gpu 2>&1 | tail -60 > gpurun_out/r2c_norm.log; tail -30 gpurun_out/r2c_norm.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
for k,v in d["kernel_breakdown"].items(): print(k, v)
PY
# ncu --set full: one steady-state instance of each hot kernel (second update of a 2-update run)
for spec in "tc_gemm_kernel:70:3" "conv_fwd_mma_kernel:70:1" "conv_bwd_mma_kernel:10:1" "row_bwd_kernel:10:1"; do
  pat=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 420 ncu --set full --import-source on --clock-control none -k regex:$pat --launch-skip $skip --launch-count $cnt \
    -o gpurun_out/r2c_ncu_$pat -f python bench.py --steps 1 --warmup 1 --no-cpu --no-env-roofline > gpurun_out/r2c_ncu_$pat.log 2>&1
  ls -la gpurun_out/r2c_ncu_$pat.ncu-rep 2>/dev/null
done
