"""NumPy restatement of the in-tree PQN arithmetic of mttga/purejaxql.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows (``purejaxql/pqn_minatar.py`` unless noted).  flax /
optax defaults relied on (third-party, restated from their published source):
``nn.LayerNorm`` (last axis, eps 1e-6, fast variance E[x^2]-E[x]^2 clamped >= 0),
``nn.BatchNorm`` (momentum 0.99, eps 1e-5), ``optax.radam`` (b1 .9, b2 .999,
eps 1e-8, threshold 5), ``optax.clip_by_global_norm``, ``optax.linear_schedule``.

Parameters are flat dicts keyed by the flax path joined with "/" (Appendix C of
SURVEY.md), e.g. ``"CNN_0/Conv_0/kernel"`` with shape [3,3,C,16] (HWIO).
"""
from __future__ import annotations

import numpy as np

from . import jax_prng as jr

F32 = np.float32
LN_EPS = 1e-6
BN_MOM = 0.99


# --------------------------------------------------------------------------- #
# parameter helpers
# --------------------------------------------------------------------------- #
def cnn_param_shapes(C: int, A: int):
    """MinAtar QNetwork tree (pqn_minatar.py:24-69), layer_norm variant."""
    return {
        "BatchNorm_0/scale": (C,), "BatchNorm_0/bias": (C,),
        "CNN_0/Conv_0/kernel": (3, 3, C, 16), "CNN_0/Conv_0/bias": (16,),
        "CNN_0/LayerNorm_0/scale": (16,), "CNN_0/LayerNorm_0/bias": (16,),
        "CNN_0/Dense_0/kernel": (1024, 128), "CNN_0/Dense_0/bias": (128,),
        "CNN_0/LayerNorm_1/scale": (128,), "CNN_0/LayerNorm_1/bias": (128,),
        "Dense_0/kernel": (128, A), "Dense_0/bias": (A,),
    }


def mlp_param_shapes(D: int, A: int, hidden: int = 256, layers: int = 2):
    """gymnax MLP QNetwork tree (pqn_gymnax.py:29-58), layer_norm variant."""
    shapes = {"BatchNorm_0/scale": (D,), "BatchNorm_0/bias": (D,)}
    fan = D
    for l in range(layers):
        shapes[f"Dense_{l}/kernel"] = (fan, hidden)
        shapes[f"Dense_{l}/bias"] = (hidden,)
        shapes[f"LayerNorm_{l}/scale"] = (hidden,)
        shapes[f"LayerNorm_{l}/bias"] = (hidden,)
        fan = hidden
    shapes[f"Dense_{layers}/kernel"] = (fan, A)
    shapes[f"Dense_{layers}/bias"] = (A,)
    return shapes


def random_params(shapes, seed=0, dtype=F32):
    """Random (not flax-identical) weights for parity tests: both sides load the
    same arrays, so only the distribution's scale matters."""
    rng = np.random.default_rng(seed)
    p = {}
    for k, shp in shapes.items():
        if k.endswith("kernel"):
            fan_in = int(np.prod(shp[:-1]))
            p[k] = (rng.standard_normal(shp) * np.sqrt(2.0 / fan_in)).astype(dtype)
        elif k.endswith("scale"):
            p[k] = (1.0 + 0.1 * rng.standard_normal(shp)).astype(dtype)
        else:
            p[k] = (0.1 * rng.standard_normal(shp)).astype(dtype)
    return p


# --------------------------------------------------------------------------- #
# layers
# --------------------------------------------------------------------------- #
def _layer_norm_fwd(x, scale, bias):
    mean = x.mean(axis=-1, keepdims=True, dtype=x.dtype)
    mean2 = (x * x).mean(axis=-1, keepdims=True, dtype=x.dtype)
    var = np.maximum(mean2 - mean * mean, 0)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(LN_EPS))
    xhat = (x - mean) * rstd
    return xhat * scale + bias, (xhat, rstd)


def _layer_norm_bwd(dy, cache, scale):
    xhat, rstd = cache
    dscale = (dy * xhat).reshape(-1, xhat.shape[-1]).sum(0)
    dbias = dy.reshape(-1, xhat.shape[-1]).sum(0)
    dxhat = dy * scale
    n = xhat.shape[-1]
    dx = rstd * (dxhat - dxhat.mean(-1, keepdims=True) - xhat * (dxhat * xhat).mean(-1, keepdims=True))
    return dx, dscale, dbias


def _im2col(x):
    """x [B,10,10,C] -> patches [B,8,8,9*C] with (di,dj,c) minor order (HWIO)."""
    B, H, W, C = x.shape
    cols = [x[:, di:di + 8, dj:dj + 8, :] for di in range(3) for dj in range(3)]
    return np.concatenate(cols, axis=-1)


def bn_batch_stats_update(batch_stats, x):
    """Dummy input BatchNorm in train mode (pqn_minatar.py:65, flax BatchNorm
    momentum 0.99): running mean/var of the raw obs over all but the last axis."""
    xr = x.reshape(-1, x.shape[-1])
    mean = xr.mean(0, dtype=x.dtype)
    var = np.maximum((xr * xr).mean(0, dtype=x.dtype) - mean * mean, 0)
    m = x.dtype.type(BN_MOM)
    one = x.dtype.type(1.0)
    return {"mean": m * batch_stats["mean"] + (one - m) * mean,
            "var": m * batch_stats["var"] + (one - m) * var}


# --------------------------------------------------------------------------- #
# MinAtar CNN Q-network (pqn_minatar.py:24-69)
# --------------------------------------------------------------------------- #
def cnn_forward(p, obs, want_cache=False):
    dt = p["CNN_0/Dense_0/kernel"].dtype
    x = obs.astype(dt) / dt.type(255.0)                                   # :66
    cols = _im2col(x)                                                      # [B,8,8,9C]
    Wc = p["CNN_0/Conv_0/kernel"].reshape(-1, 16)
    z1 = cols @ Wc + p["CNN_0/Conv_0/bias"]                                # :38-44
    y1, c1 = _layer_norm_fwd(z1, p["CNN_0/LayerNorm_0/scale"], p["CNN_0/LayerNorm_0/bias"])
    h1 = np.maximum(y1, 0).reshape(obs.shape[0], -1)                       # :46-47
    z2 = h1 @ p["CNN_0/Dense_0/kernel"] + p["CNN_0/Dense_0/bias"]          # :48
    y2, c2 = _layer_norm_fwd(z2, p["CNN_0/LayerNorm_1/scale"], p["CNN_0/LayerNorm_1/bias"])
    h2 = np.maximum(y2, 0)
    q = h2 @ p["Dense_0/kernel"] + p["Dense_0/bias"]                       # :68
    if want_cache:
        return q, (cols, c1, y1, h1, c2, y2, h2)
    return q


def cnn_loss_and_grads(p, obs, action, target):
    """``_loss_fn`` + ``value_and_grad`` (pqn_minatar.py:271-291)."""
    q, (cols, c1, y1, h1, c2, y2, h2) = cnn_forward(p, obs, want_cache=True)
    B = obs.shape[0]
    dt = q.dtype
    q_sa = q[np.arange(B), action]
    diff = q_sa - target.astype(dt)
    loss = dt.type(0.5) * np.mean(diff * diff, dtype=dt)
    dq = np.zeros_like(q)
    dq[np.arange(B), action] = diff / dt.type(B)
    g = {k: np.zeros_like(v) for k, v in p.items()}
    g["Dense_0/kernel"] = h2.T @ dq
    g["Dense_0/bias"] = dq.sum(0)
    dh2 = dq @ p["Dense_0/kernel"].T
    dy2 = dh2 * (y2 > 0)
    dz2, g["CNN_0/LayerNorm_1/scale"], g["CNN_0/LayerNorm_1/bias"] = _layer_norm_bwd(
        dy2, c2, p["CNN_0/LayerNorm_1/scale"])
    g["CNN_0/Dense_0/kernel"] = h1.T @ dz2
    g["CNN_0/Dense_0/bias"] = dz2.sum(0)
    dh1 = (dz2 @ p["CNN_0/Dense_0/kernel"].T).reshape(y1.shape)
    dy1 = dh1 * (y1 > 0)
    dz1, g["CNN_0/LayerNorm_0/scale"], g["CNN_0/LayerNorm_0/bias"] = _layer_norm_bwd(
        dy1, c1, p["CNN_0/LayerNorm_0/scale"])
    g["CNN_0/Conv_0/kernel"] = (cols.reshape(-1, cols.shape[-1]).T @ dz1.reshape(-1, 16)).reshape(
        p["CNN_0/Conv_0/kernel"].shape)
    g["CNN_0/Conv_0/bias"] = dz1.reshape(-1, 16).sum(0)
    return loss, q_sa, g


# --------------------------------------------------------------------------- #
# gymnax MLP Q-network (pqn_gymnax.py:29-58)
# --------------------------------------------------------------------------- #
def _mlp_layers(p):
    return sum(1 for k in p if k.startswith("LayerNorm_") and k.endswith("scale"))


def mlp_forward(p, obs, want_cache=False):
    L = _mlp_layers(p)
    dt = p["Dense_0/kernel"].dtype
    x = obs.astype(dt)
    cache = []
    for l in range(L):
        z = x @ p[f"Dense_{l}/kernel"] + p[f"Dense_{l}/bias"]
        y, c = _layer_norm_fwd(z, p[f"LayerNorm_{l}/scale"], p[f"LayerNorm_{l}/bias"])
        h = np.maximum(y, 0)
        cache.append((x, c, y))
        x = h
    q = x @ p[f"Dense_{L}/kernel"] + p[f"Dense_{L}/bias"]
    if want_cache:
        return q, (cache, x)
    return q


def mlp_loss_and_grads(p, obs, action, target):
    L = _mlp_layers(p)
    q, (cache, hL) = mlp_forward(p, obs, want_cache=True)
    B = obs.shape[0]
    dt = q.dtype
    q_sa = q[np.arange(B), action]
    diff = q_sa - target.astype(dt)
    loss = dt.type(0.5) * np.mean(diff * diff, dtype=dt)
    dq = np.zeros_like(q)
    dq[np.arange(B), action] = diff / dt.type(B)
    g = {k: np.zeros_like(v) for k, v in p.items()}
    g[f"Dense_{L}/kernel"] = hL.T @ dq
    g[f"Dense_{L}/bias"] = dq.sum(0)
    dh = dq @ p[f"Dense_{L}/kernel"].T
    for l in reversed(range(L)):
        x, c, y = cache[l]
        dy = dh * (y > 0)
        dz, g[f"LayerNorm_{l}/scale"], g[f"LayerNorm_{l}/bias"] = _layer_norm_bwd(
            dy, c, p[f"LayerNorm_{l}/scale"])
        g[f"Dense_{l}/kernel"] = x.T @ dz
        g[f"Dense_{l}/bias"] = dz.sum(0)
        dh = dz @ p[f"Dense_{l}/kernel"].T
    return loss, q_sa, g


# --------------------------------------------------------------------------- #
# exploration, schedules, targets
# --------------------------------------------------------------------------- #
def eps_greedy(keys, q_vals, eps):
    """``eps_greedy_exploration`` vmapped over envs (pqn_minatar.py:115-128,194-196).
    keys uint32[N,2], q_vals [N,A], eps scalar or [N]."""
    ks = jr.split(keys, 2)
    rng_a, rng_e = ks[..., 0, :], ks[..., 1, :]
    greedy = np.argmax(q_vals, axis=-1).astype(np.int32)          # first max on ties
    u = jr.uniform(rng_e, ())
    r = jr.randint(rng_a, (), 0, q_vals.shape[-1])
    return np.where(u < np.float32(eps), r, greedy).astype(np.int32)


def linear_schedule(init, end, transition_steps, count):
    """``optax.linear_schedule`` evaluated in float32 like the traced program."""
    if transition_steps <= 0:
        return F32(init)
    c = np.clip(F32(count), F32(0), F32(transition_steps))
    frac = F32(1) - c / F32(transition_steps)
    return F32(F32(init - end) * frac + F32(end))


def q_lambda_targets(reward, done, q_val, last_q, gamma, lam):
    """Q(lambda) reverse scan (pqn_minatar.py:237-260).  reward/done [T,N],
    q_val [T,N,A], last_q [N] = max_a Q(next_obs[T-1]) (unmasked)."""
    T = reward.shape[0]
    dt = reward.dtype
    gamma = dt.type(gamma)
    lam = dt.type(lam)
    one = dt.type(1)
    d = done.astype(dt)
    last_q = last_q * (one - d[-1])                                # :252
    lam_ret = reward[-1] + gamma * last_q                          # :253
    out = np.empty_like(reward)
    out[-1] = lam_ret
    next_q = last_q
    for t in range(T - 2, -1, -1):                                 # :237-250
        boot = reward[t] + gamma * (one - d[t]) * next_q
        delta = lam_ret - next_q
        lam_ret = boot + gamma * lam * delta
        lam_ret = (one - d[t]) * lam_ret + d[t] * reward[t]
        next_q = q_val[t].max(axis=-1)
        out[t] = lam_ret
    return out


# --------------------------------------------------------------------------- #
# optimizer: optax.chain(clip_by_global_norm, radam)  (pqn_minatar.py:159-162)
# --------------------------------------------------------------------------- #
def radam_scalars(count_inc, b1=0.9, b2=0.999, threshold=5.0):
    """Per-step scalars of optax.scale_by_radam for 1-based step ``count_inc``:
    (bias_corr1, bias_corr2, rect, use_rect).  float32 in optax's operation order: the traced program
    evaluates ``b2t = b2**count_inc; ro = ro_inf - 2*count_inc*b2t/(1-b2t)`` on weak-typed float32 scalars, and
    the cancellation in ``ro`` is visible at that precision for small counts."""
    f = F32
    t = f(count_inc)
    ro_inf = f(2.0) / (f(1.0) - f(b2)) - f(1.0)
    b2t = np.power(f(b2), t, dtype=F32)
    ro = ro_inf - f(2.0) * t * b2t / (f(1.0) - b2t)
    bc1 = f(1.0) - np.power(f(b1), t, dtype=F32)
    bc2 = f(1.0) - b2t
    use = bool(ro >= f(threshold))
    rect = np.sqrt((ro - f(4)) * (ro - f(2)) * ro_inf / ((ro_inf - f(4)) * (ro_inf - f(2)) * ro), dtype=F32) if use else f(0.0)
    return float(bc1), float(bc2), float(rect), use


def radam_clip_step(p, g, opt, lr, max_grad_norm, b1=0.9, b2=0.999, eps=1e-8):
    """One ``apply_gradients`` with chain(clip_by_global_norm, radam(lr)).
    opt = {"count": int, "mu": {...}, "nu": {...}}; returns new (p, opt)."""
    dt = next(iter(p.values())).dtype
    gn = np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for v in g.values()))
    if gn < max_grad_norm:
        gc = g
    else:
        gc = {k: (v / dt.type(gn)) * dt.type(max_grad_norm) for k, v in g.items()}
    count_inc = opt["count"] + 1
    bc1, bc2, rect, use = radam_scalars(count_inc, b1, b2)
    newp, mu, nu = {}, {}, {}
    for k in p:
        mu[k] = dt.type(b1) * opt["mu"][k] + dt.type(1 - b1) * gc[k]
        nu[k] = dt.type(b2) * opt["nu"][k] + dt.type(1 - b2) * gc[k] * gc[k]
        mhat = mu[k] / dt.type(bc1)
        if use:
            nhat = nu[k] / dt.type(bc2)
            upd = dt.type(rect) * mhat / (np.sqrt(nhat) + dt.type(eps))
        else:
            upd = mhat
        newp[k] = p[k] - dt.type(lr) * upd
    return newp, {"count": count_inc, "mu": mu, "nu": nu}, gn


def opt_init(p):
    return {"count": 0, "mu": {k: np.zeros_like(v) for k, v in p.items()},
            "nu": {k: np.zeros_like(v) for k, v in p.items()}}


# --------------------------------------------------------------------------- #
# rollout (pqn_minatar.py:181-219) for ONE seed
# --------------------------------------------------------------------------- #
def rollout(env, forward, params, obs, env_state, rng, num_steps, eps, rew_scale=1.0, forced_actions=None,
            tie_log=None):
    """T x _step_env.  ``rng`` is the scan's initial carry (``_rng`` of :213).
    Returns (obs_T, state_T, rng_final, transitions dict of [T,...] arrays, infos).

    ``forced_actions`` [T,N] (parity tests with eps < 1): the oracle still computes its own eps-greedy
    action, records every disagreement in ``tie_log`` as (t, env, own action, forced action, gap between the
    two Q-values involved) and then steps the env with the forced action, so that one argmax flip on a
    numerical Q tie does not desynchronise the rest of the comparison."""
    N = obs.shape[0]
    tr = {k: [] for k in ("obs", "action", "reward", "done", "next_obs", "q_val")}
    infos = {}
    for _ in range(num_steps):
        ks = jr.split(rng, 3)                                       # :183
        rng, rng_a, rng_s = ks[0], ks[1], ks[2]
        q = forward(params, obs)                                    # :184-191
        action = eps_greedy(jr.split(rng_a, N), q, eps)             # :194-196
        if forced_actions is not None:
            fa = np.asarray(forced_actions[len(tr["action"])]).astype(np.int32)
            bad = np.nonzero(fa != action)[0]
            if tie_log is not None:
                for e in bad:
                    tie_log.append((len(tr["action"]), int(e), int(action[e]), int(fa[e]),
                                    float(abs(q[e, action[e]] - q[e, fa[e]]))))
            action = fa
        new_obs, env_state, reward, done, info = env.step(jr.split(rng_s, N), env_state, action)
        tr["obs"].append(obs); tr["action"].append(action)
        tr["reward"].append((F32(rew_scale) * reward).astype(F32)); tr["done"].append(done)
        tr["next_obs"].append(new_obs); tr["q_val"].append(q)
        for k, v in info.items():
            infos.setdefault(k, []).append(v)
        obs = new_obs
    tr = {k: np.stack(v) for k, v in tr.items()}
    infos = {k: np.stack(v) for k, v in infos.items()}
    return obs, env_state, rng, tr, infos


def update_step(env, kind, params, opt, batch_stats, obs, env_state, rng, cfg, n_updates, lr_fn, forced_actions=None,
                tie_log=None):
    """One ``_update_step`` (pqn_minatar.py:176-338) for one seed, exact key chain
    of SURVEY Appendix B.  ``kind`` is "cnn" or "mlp".  Returns the new carry and
    the metrics dict."""
    fwd = cnn_forward if kind == "cnn" else mlp_forward
    lossgrad = cnn_loss_and_grads if kind == "cnn" else mlp_loss_and_grads
    T, E = cfg["NUM_STEPS"], cfg["NUM_ENVS"]
    eps = linear_schedule(cfg["EPS_START"], cfg["EPS_FINISH"],
                          cfg["EPS_DECAY"] * cfg["NUM_UPDATES_DECAY"], n_updates)
    ks = jr.split(rng, 2); rng, _rng = ks[0], ks[1]                 # :213
    obs, env_state, rng, tr, infos = rollout(env, fwd, params, obs, env_state, _rng, T, eps,
                                             cfg.get("REW_SCALE", 1), forced_actions, tie_log)
    last_q = fwd(params, tr["next_obs"][-1]).max(-1)                # :227-235
    targets = q_lambda_targets(tr["reward"], tr["done"], tr["q_val"], last_q,
                               cfg["GAMMA"], cfg["LAMBDA"])
    ks = jr.split(rng, 2); rng = ks[0]                              # :324
    losses, qvs = [], []
    flat_obs = tr["obs"].reshape((T * E,) + tr["obs"].shape[2:])
    flat_act = tr["action"].reshape(-1)
    flat_tgt = targets.reshape(-1)
    nmb = cfg["NUM_MINIBATCHES"]
    for _ in range(cfg["NUM_EPOCHS"]):
        ks = jr.split(rng, 2); rng, kperm = ks[0], ks[1]            # :309
        perm = jr.permutation_indices(kperm, T * E).reshape(nmb, -1)
        ks = jr.split(rng, 2); rng = ks[0]                          # :317
        for mb in range(nmb):
            idx = perm[mb]
            loss, q_sa, g = lossgrad(params, flat_obs[idx], flat_act[idx], flat_tgt[idx])
            batch_stats = bn_batch_stats_update(batch_stats, flat_obs[idx].astype(F32))
            params, opt, _ = radam_clip_step(params, g, opt, lr_fn(opt["count"]), cfg["MAX_GRAD_NORM"])
            losses.append(loss); qvs.append(q_sa.mean())
    metrics = {"td_loss": float(np.mean(losses)), "qvals": float(np.mean(qvs))}
    metrics.update({k: float(v.astype(np.float64).mean()) for k, v in infos.items()})
    return params, opt, batch_stats, obs, env_state, rng, metrics, tr, targets


# --------------------------------------------------------------------------- #
# greedy evaluation rollout (pqn_minatar.py:371-413 / pqn_gymnax.py:364-406) for ONE seed
# --------------------------------------------------------------------------- #
INFO_KEYS = ("returned_episode_returns", "returned_episode_lengths", "timestep", "returned_episode", "discount")


def get_test_metrics(env, forward, params, rng, num_envs, num_steps, eps_test):
    """``get_test_metrics(train_state, rng)``.  Restated with its key quirks:

    * ``rng, _rng = split(rng)`` (:396); the reset uses ``split(_rng, N)`` (:397 via vmap_reset :107-109)
      and the scan carry starts at the SAME ``_rng`` (:399-401);
    * every step: ``rng, _rng = split(rng)`` (:378); the action keys are ``split(_rng, N)`` (:388-390)
      and the env keys are ``split(_rng, N)`` again inside ``vmap_step`` (:391-393, :110-112) -- the
      same per-env keys feed ``eps_greedy_exploration`` and ``env.step``;
    * result: for every info leaf, ``nanmean(where(infos["returned_episode"], x, nan))`` over the
      whole [T, N] block (:403-412) -- NaN when no episode ended."""
    ks = jr.split(rng, 2)
    _rng = ks[1]
    obs, state = env.reset(jr.split(_rng, num_envs))
    carry = _rng
    infos = {}
    for _ in range(num_steps):
        ks = jr.split(carry, 2)
        carry, sub = ks[0], ks[1]
        q = forward(params, obs)
        env_keys = jr.split(sub, num_envs)
        action = eps_greedy(env_keys, q, eps_test)
        obs, state, reward, done, info = env.step(env_keys, state, action)
        for k, v in info.items():
            infos.setdefault(k, []).append(np.asarray(v))
    infos = {k: np.stack(v) for k, v in infos.items()}
    mask = infos["returned_episode"].astype(bool)
    out = {}
    for k, v in infos.items():
        sel = v.astype(np.float64)[mask]
        out[k] = float(sel.mean()) if sel.size else float("nan")
    return out
