"""Builds and loads the CPU logic harness (tests only; see host_harness.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_host_harness.so")
SRC = os.path.join(HERE, "host_harness.cpp")
CSRC = os.path.join(os.path.dirname(HERE), "purejaxql_b200", "csrc")


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", SO])
    return ctypes.CDLL(SO)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


ENV_IDS = {"Breakout-MinAtar": 0, "Asterix-MinAtar": 1, "SpaceInvaders-MinAtar": 2, "Freeway-MinAtar": 3,
           "CartPole-v1": 16, "Acrobot-v1": 17}


class HostEnv:
    """Drives the harness like the batched C-ABI env operator."""

    def __init__(self, name, part=0, max_steps=0):
        self.lib = load()
        self.id = ENV_IDS[name]
        self.words = self.lib.h_state_words(self.id)
        self.part = part
        self.max_steps = max_steps

    def reset(self, keys, obs_dim, default_max):
        n = keys.shape[0]
        keys = np.ascontiguousarray(keys, np.uint32)
        state = np.zeros((self.words, n), np.uint32)
        obs = np.zeros((n, obs_dim), np.float32)
        self.lib.h_env_reset(self.id, ptr(keys), ptr(state), ptr(obs), ctypes.c_int64(n),
                             self.max_steps or default_max, self.part)
        return obs, state

    def step(self, keys, state, action, obs_dim, default_max):
        n = keys.shape[0]
        keys = np.ascontiguousarray(keys, np.uint32)
        action = np.ascontiguousarray(action, np.int32)
        obs = np.zeros((n, obs_dim), np.float32)
        reward = np.zeros(n, np.float32)
        done = np.zeros(n, np.uint8)
        self.lib.h_env_step(self.id, ptr(keys), ptr(state), ptr(action), ptr(obs), ptr(reward), ptr(done),
                            ctypes.c_int64(n), self.max_steps or default_max, self.part)
        return obs, state, reward, done.astype(bool)
