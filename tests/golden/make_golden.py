"""Generates the seeded oracle trajectories under tests/golden/ (see README.md).

    python tests/golden/make_golden.py

They are produced by oracle/ (NOT by the reference: jax/gymnax are not
installable here) and pin the CUDA path and the oracle to each other.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gymnax_envs as G  # noqa: E402
from oracle import jax_prng as jr  # noqa: E402


def trajectory(name, n, steps, seed, part=False, max_steps=0):
    jr.DEFAULT_PARTITIONABLE = part
    env = G.make(name)
    if max_steps:
        env.env.core.max_steps_in_episode = max_steps
    key = jr.PRNGKey(seed)
    ks = jr.split(key, 2)
    key, kr = ks[0], ks[1]
    rkeys = jr.split(kr, n)
    obs, st = env.reset(rkeys)
    out = {"reset_keys": rkeys, "obs0": obs, "step_keys": [], "action": [], "obs": [], "reward": [], "done": [],
           "ret": [], "len": []}
    for t in range(steps):
        ks = jr.split(key, 3)
        key, ka, kst = ks[0], ks[1], ks[2]
        act = jr.randint(jr.split(ka, n), (), 0, env.num_actions)
        if name == "Freeway-MinAtar":          # mostly "up": the chicken reaches the top, cars re-randomise
            act = np.where(np.arange(n) % 4 != 0, 1, act).astype(np.int32)
        sk = jr.split(kst, n)
        obs, st, r, d, info = env.step(sk, st, act)
        out["step_keys"].append(sk); out["action"].append(act); out["obs"].append(obs)
        out["reward"].append(r); out["done"].append(d)
        out["ret"].append(info["returned_episode_returns"]); out["len"].append(info["returned_episode_lengths"])
    res = {k: (np.stack(v) if isinstance(v, list) else v) for k, v in out.items()}
    if name.endswith("MinAtar"):
        res["obs0"] = np.packbits(res["obs0"].astype(bool).reshape(n, -1), axis=-1)
        res["obs"] = np.packbits(res["obs"].astype(bool).reshape(steps, n, -1), axis=-1)
    res["final_time"] = st["time"]
    jr.DEFAULT_PARTITIONABLE = False
    if max_steps:
        type(env.env.core).max_steps_in_episode = {"Breakout-MinAtar": 1000, "Freeway-MinAtar": 2500,
                                                   "SpaceInvaders-MinAtar": 1000, "Asterix-MinAtar": 1000}.get(name, 500)
    return res


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "breakout_traj_original.npz"),
                        **trajectory("Breakout-MinAtar", 48, 300, 2024))
    np.savez_compressed(os.path.join(HERE, "breakout_traj_partitionable.npz"),
                        **trajectory("Breakout-MinAtar", 48, 120, 7, part=True))
    np.savez_compressed(os.path.join(HERE, "asterix_traj_original.npz"),
                        **trajectory("Asterix-MinAtar", 32, 400, 8))
    np.savez_compressed(os.path.join(HERE, "freeway_traj_original.npz"),
                        **trajectory("Freeway-MinAtar", 32, 200, 5))
    np.savez_compressed(os.path.join(HERE, "spaceinvaders_traj_original.npz"),
                        **trajectory("SpaceInvaders-MinAtar", 32, 400, 6))
    np.savez_compressed(os.path.join(HERE, "cartpole_traj_original.npz"), **trajectory("CartPole-v1", 32, 120, 11))
    np.savez_compressed(os.path.join(HERE, "acrobot_traj_original.npz"), **trajectory("Acrobot-v1", 32, 60, 13))
    print("golden trajectories written")
