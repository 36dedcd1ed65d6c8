#!/bin/bash
# round-2 GPU session D: locate the illegal access of the modular norm path (compute-sanitizer), norm tests one
# process per case, determinism + config-geometry tests, bench with the deterministic reductions
mkdir -p gpurun_out
bash scripts/probe_ref.sh > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_norm.py -x -q -m gpu \
  -k "loss_grad and batch_norm-False-cnn" > gpurun_out/r2d_sanitize_cnn.log 2>&1
grep -n "Invalid\|at pqn\|in pqn\|========= .*kernel\|by thread" gpurun_out/r2d_sanitize_cnn.log | head -20
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_norm.py -x -q -m gpu \
  -k "loss_grad and batch_norm-False-mlp" > gpurun_out/r2d_sanitize_mlp.log 2>&1
grep -n "Invalid\|at pqn\|in pqn\|========= .*kernel\|by thread" gpurun_out/r2d_sanitize_mlp.log | head -20
for k in "loss_grad and none-False-cnn" "loss_grad and none-True-cnn" "loss_grad and layer_norm-True-cnn" "loss_grad and batch_norm-True-cnn" \
         "loss_grad and none-False-mlp" "loss_grad and none-True-mlp" "loss_grad and layer_norm-True-mlp" "loss_grad and batch_norm-True-mlp" \
         "update_step and batch_norm-False" "update_step and layer_norm-True" "update_step and none-False" "update_step and batch_norm-True"; do
  echo "### $k"; python -m pytest tests/test_gpu_norm.py -x -q -m gpu -k "$k" 2>&1 | grep -E "passed|failed|Error|assert " | head -6
done > gpurun_out/r2d_norm_cases.log 2>&1
cat gpurun_out/r2d_norm_cases.log
python -m pytest tests/test_gpu_train.py tests/test_gpu_parity_r2.py tests/test_gpu_net.py -q -m gpu -s 2>&1 | tail -30 > gpurun_out/r2d_train.log; tail -12 gpurun_out/r2d_train.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
for k,v in list(d["kernel_breakdown"].items())[:12]: print(k, v)
PY
