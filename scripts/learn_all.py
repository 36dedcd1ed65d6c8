#!/usr/bin/env python
"""Learning evidence on the GPU for every built environment with the shipped presets (shortened)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from purejaxql_b200 import config_loader, pqn_minatar, pqn_gymnax, jaxrandom as jr

def run(mod, alg, env, steps, seeds=4, extra=()):
    c = config_loader.compose([f"+alg={alg}", f"alg.ENV_NAME={env}", f"NUM_SEEDS={seeds}", "SAVE_PATH=null",
                               f"alg.TOTAL_TIMESTEPS={steps}", f"alg.TOTAL_TIMESTEPS_DECAY={steps}", *extra])
    cfg = {**c, **c["alg"]}
    t0 = time.time()
    train = mod.make_train(cfg)
    out = train(jr.to_numpy_u32(jr.split(jr.PRNGKey(0), seeds)))
    torch.cuda.synchronize()
    dt = time.time() - t0
    m = out["metrics"]
    tst = m["test/returned_episode_returns"].cpu().numpy()
    ret = m["returned_episode_returns"].cpu().numpy()
    n = ret.shape[1]
    idx = [0, n // 4, n // 2, n - 1]
    return {"env": env, "alg": alg, "seeds": seeds, "total_timesteps": float(steps), "num_updates": n, "wall_s": round(dt, 1),
            "graph": bool(getattr(train.engine, "graph_captured", False)),
            "train_return@update": {int(i): round(float(ret[:, i].mean()), 2) for i in idx},
            "greedy_eval_return@update": {int(i): round(float(np.nanmean(tst[:, i])), 2) for i in idx}}

res = []
for env, steps in (("Asterix-MinAtar", "3e6"), ("SpaceInvaders-MinAtar", "3e6"), ("Freeway-MinAtar", "3e6")):
    res.append(run(pqn_minatar, "pqn_minatar", env, steps, extra=("alg.TEST_NUM_ENVS=32",)))
    print(json.dumps(res[-1]), flush=True)
res.append(run(pqn_gymnax, "pqn_cartpole", "CartPole-v1", "5e5", seeds=8))
print(json.dumps(res[-1]), flush=True)
res.append(run(pqn_gymnax, "pqn_cartpole", "Acrobot-v1", "5e5", seeds=8))
print(json.dumps(res[-1]), flush=True)
# BASELINE config 4 geometry: Acrobot, 65536 envs (TOTAL_TIMESTEPS overridden so that NUM_UPDATES = 20)
res.append(run(pqn_gymnax, "pqn_cartpole", "Acrobot-v1", str(65536 * 64 * 20), seeds=1,
               extra=("alg.NUM_ENVS=65536", "alg.TEST_NUM_ENVS=128")))
print(json.dumps(res[-1]), flush=True)
# round 2: the recurrent (GRU) engine with its shipped preset (pqn_rnn_cartpole.yaml, 5e5 steps, 4 seeds)
from purejaxql_b200 import pqn_rnn_gymnax
res.append(run(pqn_rnn_gymnax, "pqn_rnn_cartpole", "CartPole-v1", "5e5", seeds=4))
print(json.dumps(res[-1]), flush=True)
