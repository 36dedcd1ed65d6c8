#!/usr/bin/env python
"""Learning sanity on the GPU: PQN on Breakout-MinAtar with the shipped preset (shortened),
prints the behaviour-policy and greedy-eval returns over training."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from purejaxql_b200 import config_loader, pqn_minatar, jaxrandom as jr
steps = sys.argv[1] if len(sys.argv) > 1 else "3e6"
c = config_loader.compose(["+alg=pqn_minatar", "alg.ENV_NAME=Breakout-MinAtar", "NUM_SEEDS=8", "SAVE_PATH=null",
                           f"alg.TOTAL_TIMESTEPS={steps}", f"alg.TOTAL_TIMESTEPS_DECAY={steps}", "alg.TEST_NUM_ENVS=64"])
cfg = {**c, **c["alg"]}
t0 = time.time()
train = pqn_minatar.make_train(cfg)
out = train(jr.to_numpy_u32(jr.split(jr.PRNGKey(0), 8)))
torch.cuda.synchronize()
dt = time.time() - t0
m = out["metrics"]
ret = m["returned_episode_returns"].cpu().numpy()
tst = m["test/returned_episode_returns"].cpu().numpy()
n = ret.shape[1]
idx = [0, n // 8, n // 4, n // 2, 3 * n // 4, n - 1]
print(json.dumps({"env": "Breakout-MinAtar", "seeds": 8, "total_timesteps": float(steps), "num_updates": n, "wall_s": round(dt, 1),
                  "env_steps_per_s_incl_eval": round(8 * float(steps) / dt),
                  "train_return_mean_over_seeds@update": {int(i): round(float(ret[:, i].mean()), 3) for i in idx},
                  "greedy_eval_return@update": {int(i): round(float(np.nanmean(tst[:, i])), 3) for i in idx},
                  "td_loss_last": float(m["td_loss"][:, -1].mean())}))
