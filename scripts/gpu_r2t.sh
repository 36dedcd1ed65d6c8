#!/bin/bash
# round-2 GPU session T (1 GPU): unconditional prefetch loads in the conv kernels (ncu r2m hot spot)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_net.py tests/test_gpu_train.py -q -m gpu 2>&1 | tail -8 > gpurun_out/r2t_tests.log; tail -3 gpurun_out/r2t_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; tail -2 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], d["roofline"]["kernel"], d["roofline"]["frac"])
for k,v in list(d["kernel_breakdown"].items())[:10]: print(k,v)
PY
timeout 300 python bench.py --seeds 16 --steps 10 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2t_bench_16seeds.json 2> gpurun_out/r2t_bench_16seeds.err; tail -2 gpurun_out/r2t_bench_16seeds.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench_16seeds.json').read().strip().splitlines()[-1])
print("16 seeds: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"], "graph", d["cuda_graph"])
tot=sum(v["ms_per_update"] for v in d["kernel_breakdown"].values()); print("sum kernels (eager pass)", tot)
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k,v)
PY
