#!/bin/bash
# round-2 GPU session B: fp16-split GEMM path -- unit tests, network/train parity, bench, launch list
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2b_tc.log; tail -5 gpurun_out/r2b_tc.log
python -m pytest tests/test_gpu_net.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r2b_net.log; tail -12 gpurun_out/r2b_net.log
python -m pytest tests/test_gpu_train.py tests/test_gpu_parity_r2.py -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r2b_train.log; tail -25 gpurun_out/r2b_train.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
for k,v in d["kernel_breakdown"].items(): print(k, v)
PY
