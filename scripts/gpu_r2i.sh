#!/bin/bash
# round-2 GPU session I (1 GPU): full GPU suite, Acrobot line with the split-K weight gradient, headline line
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multi.py 2>&1 | tail -40 > gpurun_out/r2i_tests.log; tail -6 gpurun_out/r2i_tests.log
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2i_bench_acrobot.json 2> gpurun_out/r2i_bench_acrobot.err; tail -2 gpurun_out/r2i_bench_acrobot.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_acrobot.json').read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:10]: print(k,v)
PY
python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -2 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d["roofline"]["frac"], d["roofline"]["traffic"])
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k,v)
PY
