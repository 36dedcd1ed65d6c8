#!/bin/bash
# round-2 GPU session G (1 GPU): norm failure detail, MLP tensor-core path tests, Acrobot bench line, headline bench
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
python -m pytest tests/test_gpu_norm.py -x -q -m gpu -k "loss_grad and batch_norm-False-cnn" 2>&1 | grep -E "^E  |assert|passed|failed" | head -20 > gpurun_out/r2g_norm_detail.log; cat gpurun_out/r2g_norm_detail.log | cut -c1-300
python -m pytest tests/test_gpu_net.py tests/test_gpu_norm.py tests/test_gpu_train.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2g_tests.log; tail -6 gpurun_out/r2g_tests.log
python bench.py --config acrobot65536 --steps 4 --warmup 3 > gpurun_out/r2g_bench_acrobot.json 2> gpurun_out/r2g_bench_acrobot.err; tail -3 gpurun_out/r2g_bench_acrobot.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_acrobot.json').read().strip().splitlines()[-1])
print(d["metric"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:10]: print(k,v)
PY
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d.get("cpu_baseline",{}).get("value"))
PY
