#!/usr/bin/env python
"""Micro-benchmark of the HBM-bound environment kernels (BASELINE
configs 3/4 style): standalone pqn_env_step with fp32 observations and the fused
rollout step with bit-packed observations, timed with CUDA events, inputs larger
than L2.  Prints one JSON object; used for profiles/ and DESIGN.md."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from purejaxql_b200 import _lib, envs, jaxrandom as jr  # noqa: E402


def time_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device("cuda:0")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    out = {"hbm_peak_gbs": hbm, "peak_source": "measured" if peaks else "fallback"}
    L = _lib.lib()
    for name, n in (("Breakout-MinAtar", 1 << 20), ("CartPole-v1", 1 << 24), ("Acrobot-v1", 1 << 24)):
        env, params = envs.make(name)
        keys = jr.split(jr.PRNGKey(0, dev), n)
        obs, st = env.reset(keys, params)
        act = torch.randint(0, env.num_actions, (n,), dtype=torch.int32, device=dev)
        o = torch.empty((n, env.obs_dim), dtype=torch.float32, device=dev)
        r = torch.empty(n, device=dev); d = torch.empty(n, dtype=torch.uint8, device=dev)
        i0 = torch.empty(n, device=dev); i1 = torch.empty(n, device=dev)
        i2 = torch.empty(n, dtype=torch.int32, device=dev); i3 = torch.empty(n, dtype=torch.int32, device=dev)

        def step():
            _lib.check(L.pqn_env_step(env.env_id, _lib.p(keys), _lib.p(st), _lib.p(act), _lib.p(o), _lib.p(r), _lib.p(d),
                                      _lib.p(i0), _lib.p(i1), _lib.p(i2), _lib.p(i3), n, 0, 0, _lib.stream_ptr()))
        ms = time_ms(step)
        sw = env.state_words * 4
        bytes_per = (sw + 4 + 8) + (sw + env.obs_dim * 4 + 4 + 1 + 16)      # read state/action/key ; write state/obs/reward/done/info
        gbs = bytes_per * n / (ms * 1e-3) / 1e9
        out[name] = {"kernel": "env_step_kernel (standalone op, fp32 obs)", "envs": n, "ms": round(ms, 4),
                     "alg_bytes_per_env_step": bytes_per, "env_steps_per_s": n / (ms * 1e-3), "achieved_gbs": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / hbm, 4)}
    # fused rollout step (packed obs) at the bench geometry: S=128, E=4096
    S, E, T = 128, 4096, 32
    env, params = envs.make("Breakout-MinAtar")
    n = S * E
    keys = jr.split(jr.PRNGKey(1, dev), n)
    _, st = env.reset(keys, params)
    q = torch.randn((n, 3), device=dev)
    eps = torch.tensor([0.5], device=dev)
    sk = jr.split(jr.PRNGKey(2, dev), S * 2).view(S, 2, 2).contiguous()
    obsn = torch.zeros((S, 2, E, 16), dtype=torch.int32, device=dev)
    a = torch.zeros((S, E), dtype=torch.int32, device=dev); rw = torch.zeros((S, E), device=dev)
    dn = torch.zeros((S, E), dtype=torch.uint8, device=dev); mq = torch.zeros((S, E), device=dev)
    sums = torch.zeros((S, 5), dtype=torch.float64, device=dev)

    def fused():
        _lib.check(L.pqn_rollout_act_step(env.env_id, _lib.p(sk), _lib.p(q), _lib.p(eps), _lib.p(st), _lib.raw(obsn[:, 1]),
                                          2 * E, _lib.p(a), _lib.p(rw), _lib.p(dn), _lib.p(mq), E, _lib.p(sums), 0, S, E, 0, 0,
                                          1000, 1.0, 0, _lib.stream_ptr()))
    ms = time_ms(fused, iters=50)
    bytes_per = (44 + 12) + (44 + 64 + 4 + 4 + 1 + 4)                    # read state + q ; write state, packed obs, a, r, done, maxq
    gbs = bytes_per * n / (ms * 1e-3) / 1e9
    out["rollout_act_step(Breakout, packed obs)"] = {
        "envs": n, "ms": round(ms, 4), "alg_bytes_per_env_step": bytes_per, "env_steps_per_s": n / (ms * 1e-3),
        "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 4),
        "note": "17 threefry blocks + game logic per env-step; obs bit-packed (64 B) so the kernel is integer-issue bound, not HBM bound"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
