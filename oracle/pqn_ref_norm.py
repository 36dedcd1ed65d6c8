"""Normalisation variants of the PQN Q-networks: ``NORM_TYPE`` in {layer_norm, batch_norm, none} and
``NORM_INPUT`` (purejaxql/pqn_minatar.py:24-69, purejaxql/pqn_gymnax.py:29-58).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  This module restates the configurations the CUDA path does not
build yet (the shipped YAMLs of the hot path use layer_norm / NORM_INPUT=False, which
``oracle/pqn_ref.py`` covers and the kernels are checked against).  It is the oracle-first step of that row: the
arithmetic is pinned here by finite differences and by exact agreement with ``pqn_ref`` on the shared configuration,
so that the kernels of the next round have something to be compared with.

flax semantics restated (third party, from the published source of flax.linen.normalization):
``nn.BatchNorm(use_running_average=not train)``: reduction over every axis but the last; momentum 0.99, eps 1e-5,
fast variance ``max(E[x^2] - E[x]^2, 0)``; in train mode the batch statistics normalise the output and the running
statistics become ``m * ra + (1 - m) * batch``; in eval mode the running statistics normalise.
Module auto-naming: inside ``CNN`` the two ``normalize`` calls become ``CNN_0/BatchNorm_0`` and ``CNN_0/BatchNorm_1``
(replacing ``LayerNorm_0/1``); in the MLP the input BatchNorm is ``BatchNorm_0`` and the hidden ones follow as
``BatchNorm_1..L`` because they share the module's counter (with layer_norm they are ``LayerNorm_0..L-1``).
"""
from __future__ import annotations

import numpy as np

from .pqn_ref import BN_MOM, _im2col, _layer_norm_bwd, _layer_norm_fwd

BN_EPS = 1e-5


# --------------------------------------------------------------------------- #
# BatchNorm
# --------------------------------------------------------------------------- #
def batch_norm_fwd(x, scale, bias, stats, train):
    """-> (y, cache, new_stats).  ``stats`` = {"mean": [F], "var": [F]} running statistics."""
    dt = x.dtype
    xr = x.reshape(-1, x.shape[-1])
    if train:
        mean = xr.mean(0, dtype=dt)
        var = np.maximum((xr * xr).mean(0, dtype=dt) - mean * mean, 0)
        m = dt.type(BN_MOM)
        new_stats = {"mean": m * stats["mean"] + (dt.type(1) - m) * mean,
                     "var": m * stats["var"] + (dt.type(1) - m) * var}
    else:
        mean, var, new_stats = stats["mean"].astype(dt), stats["var"].astype(dt), stats
    rstd = dt.type(1) / np.sqrt(var + dt.type(BN_EPS))
    xhat = (x - mean) * rstd
    return xhat * scale + bias, (xhat, rstd, train), new_stats


def batch_norm_bwd(dy, cache, scale):
    """-> (dx, dscale, dbias); in eval mode the statistics are constants."""
    xhat, rstd, train = cache
    F = xhat.shape[-1]
    dscale = (dy * xhat).reshape(-1, F).sum(0)
    dbias = dy.reshape(-1, F).sum(0)
    dxhat = dy * scale
    if not train:
        return dxhat * rstd, dscale, dbias
    red = tuple(range(dy.ndim - 1))
    dx = rstd * (dxhat - dxhat.mean(red, keepdims=True) - xhat * (dxhat * xhat).mean(red, keepdims=True))
    return dx, dscale, dbias


def _norm_fwd(kind, x, p, name, stats, train):
    """Apply ``normalize`` number `name` ("CNN_0/%s_0" style prefix handled by the caller)."""
    if kind == "layer_norm":
        y, c = _layer_norm_fwd(x, p[name + "/scale"], p[name + "/bias"])
        return y, ("ln", c), None
    if kind == "batch_norm":
        y, c, ns = batch_norm_fwd(x, p[name + "/scale"], p[name + "/bias"], stats[name], train)
        return y, ("bn", c), ns
    return x, ("id", None), None


def _norm_bwd(dy, cache, scale):
    tag, c = cache
    if tag == "ln":
        return _layer_norm_bwd(dy, c, scale)
    if tag == "bn":
        return batch_norm_bwd(dy, c, scale)
    return dy, None, None


def _norm_name(kind, prefix, idx, bn_offset=0):
    if kind == "layer_norm":
        return f"{prefix}LayerNorm_{idx}"
    if kind == "batch_norm":
        return f"{prefix}BatchNorm_{idx + bn_offset}"
    return None


# --------------------------------------------------------------------------- #
# parameter / batch_stats trees
# --------------------------------------------------------------------------- #
def cnn_param_shapes(C, A, norm_type="layer_norm"):
    s = {"BatchNorm_0/scale": (C,), "BatchNorm_0/bias": (C,),
         "CNN_0/Conv_0/kernel": (3, 3, C, 16), "CNN_0/Conv_0/bias": (16,),
         "CNN_0/Dense_0/kernel": (1024, 128), "CNN_0/Dense_0/bias": (128,),
         "Dense_0/kernel": (128, A), "Dense_0/bias": (A,)}
    for i, f in ((0, 16), (1, 128)):
        n = _norm_name(norm_type, "CNN_0/", i)
        if n:
            s[n + "/scale"], s[n + "/bias"] = (f,), (f,)
    return s


def cnn_batch_stats(C, norm_type="layer_norm", dtype=np.float32):
    feats = {"BatchNorm_0": C}
    if norm_type == "batch_norm":
        feats.update({"CNN_0/BatchNorm_0": 16, "CNN_0/BatchNorm_1": 128})
    return {k: {"mean": np.zeros(f, dtype), "var": np.ones(f, dtype)} for k, f in feats.items()}


def mlp_param_shapes(D, A, hidden=256, layers=2, norm_type="layer_norm"):
    s = {"BatchNorm_0/scale": (D,), "BatchNorm_0/bias": (D,)}
    d_in = D
    for l in range(layers):
        s[f"Dense_{l}/kernel"], s[f"Dense_{l}/bias"] = (d_in, hidden), (hidden,)
        n = _norm_name(norm_type, "", l, bn_offset=1)
        if n:
            s[n + "/scale"], s[n + "/bias"] = (hidden,), (hidden,)
        d_in = hidden
    s[f"Dense_{layers}/kernel"], s[f"Dense_{layers}/bias"] = (hidden, A), (A,)
    return s


def mlp_batch_stats(D, hidden=256, layers=2, norm_type="layer_norm", dtype=np.float32):
    feats = {"BatchNorm_0": D}
    if norm_type == "batch_norm":
        feats.update({f"BatchNorm_{l + 1}": hidden for l in range(layers)})
    return {k: {"mean": np.zeros(f, dtype), "var": np.ones(f, dtype)} for k, f in feats.items()}


# --------------------------------------------------------------------------- #
# MinAtar CNN Q-network
# --------------------------------------------------------------------------- #
def cnn_forward(p, batch_stats, obs, train, norm_type="layer_norm", norm_input=False, want_cache=False):
    """-> q (and cache), new batch_stats (== batch_stats in eval mode).  pqn_minatar.py:24-69."""
    dt = p["CNN_0/Dense_0/kernel"].dtype
    new_stats = dict(batch_stats)
    x_in = obs.astype(dt)
    y0, c0, new_stats["BatchNorm_0"] = batch_norm_fwd(x_in, p["BatchNorm_0/scale"], p["BatchNorm_0/bias"],
                                                      batch_stats["BatchNorm_0"], train)          # :61-65
    x = y0 if norm_input else x_in / dt.type(255.0)                                                # :62,66
    cols = _im2col(x)
    z1 = cols @ p["CNN_0/Conv_0/kernel"].reshape(-1, 16) + p["CNN_0/Conv_0/bias"]                  # :38-44
    n0, n1 = _norm_name(norm_type, "CNN_0/", 0), _norm_name(norm_type, "CNN_0/", 1)
    y1, c1, s1 = _norm_fwd(norm_type, z1, p, n0, batch_stats, train)                               # :45
    h1 = np.maximum(y1, 0).reshape(obs.shape[0], -1)                                               # :46-47
    z2 = h1 @ p["CNN_0/Dense_0/kernel"] + p["CNN_0/Dense_0/bias"]                                  # :48
    y2, c2, s2 = _norm_fwd(norm_type, z2, p, n1, batch_stats, train)                               # :49
    h2 = np.maximum(y2, 0)
    q = h2 @ p["Dense_0/kernel"] + p["Dense_0/bias"]                                               # :68
    if s1 is not None:
        new_stats[n0], new_stats[n1] = s1, s2
    if want_cache:
        return q, (c0, cols, c1, y1, h1, c2, y2, h2), new_stats
    return q, new_stats


def cnn_loss_and_grads(p, batch_stats, obs, action, target, norm_type="layer_norm", norm_input=False):
    """``_loss_fn`` (train=True, mutable batch_stats) + ``value_and_grad`` (pqn_minatar.py:271-296).
    -> loss, q_sa, grads, new batch_stats."""
    q, (c0, cols, c1, y1, h1, c2, y2, h2), new_stats = cnn_forward(p, batch_stats, obs, True, norm_type, norm_input, True)
    B, dt = obs.shape[0], q.dtype
    q_sa = q[np.arange(B), action]
    diff = q_sa - target.astype(dt)
    loss = dt.type(0.5) * np.mean(diff * diff, dtype=dt)
    dq = np.zeros_like(q)
    dq[np.arange(B), action] = diff / dt.type(B)
    g = {k: np.zeros_like(v) for k, v in p.items()}
    n0, n1 = _norm_name(norm_type, "CNN_0/", 0), _norm_name(norm_type, "CNN_0/", 1)
    g["Dense_0/kernel"], g["Dense_0/bias"] = h2.T @ dq, dq.sum(0)
    dy2 = (dq @ p["Dense_0/kernel"].T) * (y2 > 0)
    dz2, ds, db = _norm_bwd(dy2, c2, p[n1 + "/scale"] if n1 else None)
    if n1:
        g[n1 + "/scale"], g[n1 + "/bias"] = ds, db
    g["CNN_0/Dense_0/kernel"], g["CNN_0/Dense_0/bias"] = h1.T @ dz2, dz2.sum(0)
    dy1 = (dz2 @ p["CNN_0/Dense_0/kernel"].T).reshape(y1.shape) * (y1 > 0)
    dz1, ds, db = _norm_bwd(dy1, c1, p[n0 + "/scale"] if n0 else None)
    if n0:
        g[n0 + "/scale"], g[n0 + "/bias"] = ds, db
    g["CNN_0/Conv_0/kernel"] = (cols.reshape(-1, cols.shape[-1]).T @ dz1.reshape(-1, 16)).reshape(
        p["CNN_0/Conv_0/kernel"].shape)
    g["CNN_0/Conv_0/bias"] = dz1.reshape(-1, 16).sum(0)
    if norm_input:  # the input BatchNorm is on the path: gradient flows into its scale / bias
        Wc = p["CNN_0/Conv_0/kernel"]
        dcols = dz1 @ Wc.reshape(-1, 16).T                                  # [B,8,8,9C]
        C = obs.shape[-1]
        dx = np.zeros(obs.shape, dt)
        for k, (di, dj) in enumerate((a, b) for a in range(3) for b in range(3)):
            dx[:, di:di + 8, dj:dj + 8, :] += dcols[..., k * C:(k + 1) * C]
        _, g["BatchNorm_0/scale"], g["BatchNorm_0/bias"] = batch_norm_bwd(dx, c0, p["BatchNorm_0/scale"])
    return loss, q_sa, g, new_stats


# --------------------------------------------------------------------------- #
# gymnax MLP Q-network
# --------------------------------------------------------------------------- #
def _mlp_layers(p):
    return sum(1 for k in p if k.startswith("Dense_") and k.endswith("kernel")) - 1


def mlp_forward(p, batch_stats, obs, train, norm_type="layer_norm", norm_input=False, want_cache=False):
    """pqn_gymnax.py:29-58 (no /255)."""
    dt = p["Dense_0/kernel"].dtype
    new_stats = dict(batch_stats)
    x_in = obs.astype(dt)
    y0, c0, new_stats["BatchNorm_0"] = batch_norm_fwd(x_in, p["BatchNorm_0/scale"], p["BatchNorm_0/bias"],
                                                      batch_stats["BatchNorm_0"], train)          # :39-43
    x = y0 if norm_input else x_in
    caches = []
    for l in range(_mlp_layers(p)):
        z = x @ p[f"Dense_{l}/kernel"] + p[f"Dense_{l}/bias"]
        name = _norm_name(norm_type, "", l, bn_offset=1)
        y, c, s = _norm_fwd(norm_type, z, p, name, batch_stats, train)
        if s is not None:
            new_stats[name] = s
        h = np.maximum(y, 0)
        caches.append((x, c, y, name))
        x = h
    L = _mlp_layers(p)
    q = x @ p[f"Dense_{L}/kernel"] + p[f"Dense_{L}/bias"]
    if want_cache:
        return q, (c0, caches, x), new_stats
    return q, new_stats


def mlp_loss_and_grads(p, batch_stats, obs, action, target, norm_type="layer_norm", norm_input=False):
    q, (c0, caches, h_last), new_stats = mlp_forward(p, batch_stats, obs, True, norm_type, norm_input, True)
    B, dt = obs.shape[0], q.dtype
    L = _mlp_layers(p)
    q_sa = q[np.arange(B), action]
    diff = q_sa - target.astype(dt)
    loss = dt.type(0.5) * np.mean(diff * diff, dtype=dt)
    dq = np.zeros_like(q)
    dq[np.arange(B), action] = diff / dt.type(B)
    g = {k: np.zeros_like(v) for k, v in p.items()}
    g[f"Dense_{L}/kernel"], g[f"Dense_{L}/bias"] = h_last.T @ dq, dq.sum(0)
    dh = dq @ p[f"Dense_{L}/kernel"].T
    for l in reversed(range(L)):
        x, c, y, name = caches[l]
        dy = dh * (y > 0)
        dz, ds, db = _norm_bwd(dy, c, p[name + "/scale"] if name else None)
        if name:
            g[name + "/scale"], g[name + "/bias"] = ds, db
        g[f"Dense_{l}/kernel"], g[f"Dense_{l}/bias"] = x.T @ dz, dz.sum(0)
        dh = dz @ p[f"Dense_{l}/kernel"].T
    if norm_input:
        _, g["BatchNorm_0/scale"], g["BatchNorm_0/bias"] = batch_norm_bwd(dh, c0, p["BatchNorm_0/scale"])
    return loss, q_sa, g, new_stats
