"""CPU checks of the per-env device logic (compiled for the host by
tests/_harness.py) against the oracle: Breakout bit-exact, classic control to
fp32 tolerance, PRNG helpers bit-exact, eps-greedy and Q(lambda)."""
import ctypes

import numpy as np
import pytest
import torch

import _harness
from oracle import gymnax_envs as G
from oracle import jax_prng as jr
from oracle import pqn_ref as R
from purejaxql_b200 import envs as E

ptr = _harness.ptr


@pytest.mark.parametrize("part", [0, 1])
def test_prng_helpers_match_oracle(part):
    lib = _harness.load()
    keys = np.stack([jr.PRNGKey(s) for s in (0, 1, 42, 2 ** 31 + 5)]).astype(np.uint32)
    for num in (2, 3, 5, 128):
        out = np.zeros((4, num, 2), np.uint32)
        lib.h_split(ptr(keys), ctypes.c_int64(4), num, ptr(out), part)
        assert np.array_equal(out, jr.split(keys, num, partitionable=bool(part)))
    o3 = np.zeros(6, np.uint32)
    lib.h_split3(ptr(keys[2]), ptr(o3), part)
    assert np.array_equal(o3.reshape(3, 2), jr.split(keys[2], 3, partitionable=bool(part)))
    for ln in (1, 2, 3, 4, 7, 100):
        out = np.zeros(ln, np.uint32)
        lib.h_bits(ptr(keys[1]), ctypes.c_int64(ln), ptr(out), part)
        assert np.array_equal(out, jr.random_bits(keys[1], (ln,), partitionable=bool(part)))


def _rollout_compare(name, steps, n, part, exact, max_steps=0, atol=0.0):
    jr.DEFAULT_PARTITIONABLE = bool(part)
    try:
        env = G.make(name)
        core = env.env.core
        if max_steps:
            core.max_steps_in_episode = max_steps
        h = _harness.HostEnv(name, part=part, max_steps=max_steps)
        obs_dim = int(np.prod(core.obs_shape))
        dmax = core.max_steps_in_episode
        key = jr.PRNGKey(123)
        ks = jr.split(key, 2)
        key, kr = ks[0], ks[1]
        rkeys = jr.split(kr, n)
        o_obs, o_st = env.reset(rkeys)
        h_obs, h_st = h.reset(rkeys, obs_dim, dmax)
        assert np.allclose(h_obs, o_obs.reshape(n, -1), atol=atol, rtol=0)
        for t in range(steps):
            ks = jr.split(key, 3)
            key, ka, kst = ks[0], ks[1], ks[2]
            act = jr.randint(jr.split(ka, n), (), 0, env.num_actions)
            if name == "Freeway-MinAtar":      # mostly "up" so that the chicken reaches the top (cars re-randomise)
                act = np.where(np.arange(n) % 4 != 0, 1, act).astype(np.int32)
            skeys = jr.split(kst, n)
            if not exact:  # teacher-force the oracle state so fp32 drift cannot accumulate
                h_st = E.fields_to_state(name, {k: torch.from_numpy(np.ascontiguousarray(v))
                                                for k, v in o_st.items()}).numpy().view(np.uint32).copy()
            o_obs, o_st, o_r, o_d, o_info = env.step(skeys, o_st, act)
            h_obs, h_st, h_r, h_d = h.step(skeys, h_st, act, obs_dim, dmax)
            assert np.array_equal(h_d, o_d), (name, t)
            assert np.allclose(h_r, o_r, atol=atol, rtol=0), (name, t)
            assert np.allclose(h_obs, o_obs.reshape(n, -1), atol=atol, rtol=0), (name, t)
            f = E.state_to_fields(name, torch.from_numpy(h_st.view(np.int32)))
            for k, v in o_st.items():
                hv = f[k].numpy()
                if exact or v.dtype.kind in "ib":
                    assert np.array_equal(hv.astype(v.dtype), v), (name, t, k)
                else:
                    assert np.allclose(hv, v, atol=atol, rtol=0), (name, t, k)
    finally:
        jr.DEFAULT_PARTITIONABLE = False
        G.Breakout.max_steps_in_episode = 1000
        G.Freeway.max_steps_in_episode = 2500
        G.SpaceInvaders.max_steps_in_episode = 1000
        G.Asterix.max_steps_in_episode = 1000


@pytest.mark.parametrize("part", [0, 1])
def test_asterix_logic_bit_exact(part):
    _rollout_compare("Asterix-MinAtar", steps=900, n=96, part=part, exact=True)


def test_space_invaders_logic_bit_exact():
    _rollout_compare("SpaceInvaders-MinAtar", steps=700, n=96, part=0, exact=True)


def test_space_invaders_time_limit():
    _rollout_compare("SpaceInvaders-MinAtar", steps=40, n=32, part=0, exact=True, max_steps=11)


@pytest.mark.parametrize("part", [0, 1])
def test_freeway_logic_bit_exact(part):
    _rollout_compare("Freeway-MinAtar", steps=300, n=96, part=part, exact=True)


def test_freeway_time_limit():
    _rollout_compare("Freeway-MinAtar", steps=30, n=32, part=0, exact=True, max_steps=9)


@pytest.mark.parametrize("part", [0, 1])
def test_breakout_logic_bit_exact(part):
    _rollout_compare("Breakout-MinAtar", steps=400, n=192, part=part, exact=True)


def test_breakout_time_limit_truncation():
    _rollout_compare("Breakout-MinAtar", steps=40, n=64, part=0, exact=True, max_steps=7)


def test_cartpole_logic():
    _rollout_compare("CartPole-v1", steps=300, n=128, part=0, exact=False, atol=2e-6)


def test_acrobot_logic():
    _rollout_compare("Acrobot-v1", steps=120, n=128, part=0, exact=False, atol=2e-5)


def test_fields_roundtrip():
    env = G.make("Breakout-MinAtar")
    _, st = env.reset(jr.split(jr.PRNGKey(5), 33))
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in st.items()}
    words = E.fields_to_state("Breakout-MinAtar", t)
    back = E.state_to_fields("Breakout-MinAtar", words)
    for k, v in st.items():
        assert np.array_equal(back[k].numpy().astype(v.dtype), v), k


def test_eps_greedy_and_qlambda_logic():
    lib = _harness.load()
    rng = np.random.default_rng(0)
    n, A = 500, 3
    keys = jr.split(jr.PRNGKey(9), n)
    q = rng.standard_normal((n, A)).astype(np.float32)
    q[::7, 1] = q[::7, 0]  # ties -> first max
    for eps in (0.0, 0.3, 1.0):
        act = np.zeros(n, np.int32)
        mq = np.zeros(n, np.float32)
        lib.h_eps_greedy(ptr(keys), ptr(q), ctypes.c_float(eps), ptr(act), ptr(mq), ctypes.c_int64(n), A, 0)
        assert np.array_equal(act, R.eps_greedy(keys, q, eps))
        assert np.array_equal(mq, q.max(-1))
    T = 9
    r = rng.standard_normal((T, n)).astype(np.float32)
    d = (rng.random((T, n)) < 0.2)
    qv = rng.standard_normal((T, n, A)).astype(np.float32)
    ql = rng.standard_normal((n, A)).astype(np.float32)
    tg = np.zeros((T, n), np.float32)
    mqv = np.ascontiguousarray(qv.max(-1))
    d8 = d.astype(np.uint8)
    lib.h_qlambda(ptr(r), ptr(d8), ptr(mqv), ptr(ql), ptr(tg), T, ctypes.c_int64(n), A,
                  ctypes.c_float(0.99), ctypes.c_float(0.65))
    ref = R.q_lambda_targets(r, d, qv, ql.max(-1), 0.99, 0.65)
    assert np.allclose(tg, ref, atol=1e-6, rtol=0)
