#!/bin/bash
# round-2 GPU session F: GRU tests, full suite, headline bench (conv16 register fix), configs[2]/[3] bench lines
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python -m pytest tests/test_gpu_rnn.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2f_rnn.log; tail -25 gpurun_out/r2f_rnn.log
python -m pytest tests -q -m gpu --deselect tests/test_gpu_rnn.py 2>&1 | tail -25 > gpurun_out/r2f_suite.log; tail -8 gpurun_out/r2f_suite.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:10]: print(k, v)
PY
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2f_bench_acrobot.json 2> gpurun_out/r2f_bench_acrobot.err; tail -c 900 gpurun_out/r2f_bench_acrobot.json; tail -3 gpurun_out/r2f_bench_acrobot.err
python bench.py --config minatar5 --steps 5 --warmup 3 > gpurun_out/r2f_bench_minatar5.json 2> gpurun_out/r2f_bench_minatar5.err; cut -c1-330 gpurun_out/r2f_bench_minatar5.json; tail -3 gpurun_out/r2f_bench_minatar5.err
