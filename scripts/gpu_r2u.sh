#!/bin/bash
# round-2 GPU session U (1 GPU): HBM write/read yardsticks, GRU engine under CUDA-graph replay
mkdir -p gpurun_out
python scripts/bench_hbm_rw.py > gpurun_out/r2u_hbm_rw.json 2> gpurun_out/r2u_hbm_rw.err; cat gpurun_out/r2u_hbm_rw.json; tail -2 gpurun_out/r2u_hbm_rw.err
timeout 900 python -m pytest tests/test_gpu_rnn.py -q -m gpu 2>&1 | tail -8 > gpurun_out/r2u_rnn.log; tail -3 gpurun_out/r2u_rnn.log
timeout 600 python - > gpurun_out/r2u_rnn_learn.json 2> gpurun_out/r2u_rnn_learn.err <<'PY'
import json, time, sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from purejaxql_b200 import config_loader, pqn_rnn_gymnax, jaxrandom as jr
c = config_loader.compose(["+alg=pqn_rnn_cartpole", "NUM_SEEDS=4", "SAVE_PATH=null"])
cfg = {**c, **c["alg"]}
t0 = time.time()
train = pqn_rnn_gymnax.make_train(cfg)
out = train(jr.to_numpy_u32(jr.split(jr.PRNGKey(0), 4)))
torch.cuda.synchronize()
dt = time.time() - t0
m = out["metrics"]; ret = m["returned_episode_returns"].cpu().numpy(); tst = m["test/returned_episode_returns"].cpu().numpy()
n = ret.shape[1]
print(json.dumps({"env": "CartPole-v1", "alg": "pqn_rnn_cartpole", "seeds": 4, "num_updates": n, "wall_s": round(dt, 1), "graph": bool(train.engine.graph_captured),
                  "train_return@update": {int(i): round(float(ret[:, i].mean()), 2) for i in (0, n // 4, n // 2, n - 1)},
                  "greedy_eval_return@update": {int(i): round(float(np.nanmean(tst[:, i])), 2) for i in (0, n // 4, n // 2, n - 1)}}))
PY
cat gpurun_out/r2u_rnn_learn.json; tail -3 gpurun_out/r2u_rnn_learn.err
