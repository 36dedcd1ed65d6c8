"""NumPy restatement of the recurrent PQN network and loss of purejaxql/pqn_rnn_gymnax.py (GRU variant).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Oracle-first groundwork for SURVEY.md section 8(f) row 4: no CUDA
kernel exists or is claimed for this path yet.  Restated here: ``RNNQNetwork`` (``:57-105``: dummy input BatchNorm,
``NUM_LAYERS`` x {Dense -> LayerNorm -> ReLU}, one-hot last action appended, ``ScannedRNN`` = GRU cell scanned over
time with the carry reset to zeros where ``last_done`` is set (``:25-54``), Dense Q-head), the in-loss Q(lambda)
targets (``:295-323``) and ``_loss_fn`` (``:330-360``), with the analytic backward (BPTT) that ``value_and_grad``
computes.  Shipped configuration only: NORM_TYPE=layer_norm, NORM_INPUT=False.

flax.linen.GRUCell semantics restated (third party, published source):
    r = sigmoid(x W_ir + b_ir + h W_hr)        z = sigmoid(x W_iz + b_iz + h W_hz)
    n = tanh(x W_in + b_in + r * (h W_hn + b_hn))            h' = (1 - z) * n + z * h
(input denses carry a bias, the recurrent ones do not, except ``hn``); parameter paths
``ScannedRNN_0/GRUCell_0/{ir,iz,in,hr,hz,hn}/{kernel,bias}``.
"""
from __future__ import annotations

import numpy as np

from .pqn_ref import _layer_norm_bwd, _layer_norm_fwd

G = "ScannedRNN_0/GRUCell_0/"


def rnn_param_shapes(D, A, hidden=128, layers=2):
    """Parameter tree of RNNQNetwork (:57-105) for obs size D, A actions."""
    s = {"BatchNorm_0/scale": (D,), "BatchNorm_0/bias": (D,)}
    d_in = D
    for l in range(layers):
        s[f"Dense_{l}/kernel"], s[f"Dense_{l}/bias"] = (d_in, hidden), (hidden,)
        s[f"LayerNorm_{l}/scale"], s[f"LayerNorm_{l}/bias"] = (hidden,), (hidden,)
        d_in = hidden
    for g in ("ir", "iz", "in"):
        s[G + g + "/kernel"], s[G + g + "/bias"] = (hidden + A, hidden), (hidden,)
    for g in ("hr", "hz"):
        s[G + g + "/kernel"] = (hidden, hidden)
    s[G + "hn/kernel"], s[G + "hn/bias"] = (hidden, hidden), (hidden,)
    s[f"Dense_{layers}/kernel"], s[f"Dense_{layers}/bias"] = (hidden, A), (A,)
    return s


def _layers(p):
    return sum(1 for k in p if k.startswith("LayerNorm_") and k.endswith("scale"))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def rnn_forward(p, hs, obs, last_done, last_action, want_cache=False):
    """``network.apply(params, hs, obs, done, last_action)`` (:57-101).
    hs [B,H]; obs [T,B,D]; last_done [T,B] bool; last_action [T,B] int -> (new_hs [B,H], q [T,B,A])."""
    dt = p["Dense_0/kernel"].dtype
    T, B = obs.shape[:2]
    L = _layers(p)
    A = p[f"Dense_{L}/kernel"].shape[1]
    x = obs.astype(dt)                                       # dummy BatchNorm output is discarded (:75-76)
    trunk = []
    for l in range(L):                                       # :78-81
        z = x @ p[f"Dense_{l}/kernel"] + p[f"Dense_{l}/bias"]
        y, c = _layer_norm_fwd(z, p[f"LayerNorm_{l}/scale"], p[f"LayerNorm_{l}/bias"])
        trunk.append((x, c, y))
        x = np.maximum(y, 0)
    onehot = np.zeros((T, B, A), dt)
    np.put_along_axis(onehot, np.asarray(last_action, np.int64)[..., None], 1.0, axis=-1)
    xin = np.concatenate([x, onehot], axis=-1)               # :84-85
    h = hs.astype(dt)
    steps, ys = [], []
    for t in range(T):                                       # ScannedRNN (:35-46)
        h0 = np.where(np.asarray(last_done[t], bool)[:, None], dt.type(0), h)
        a_r = xin[t] @ p[G + "ir/kernel"] + p[G + "ir/bias"] + h0 @ p[G + "hr/kernel"]
        a_z = xin[t] @ p[G + "iz/kernel"] + p[G + "iz/bias"] + h0 @ p[G + "hz/kernel"]
        r, zg = _sigmoid(a_r), _sigmoid(a_z)
        hn = h0 @ p[G + "hn/kernel"] + p[G + "hn/bias"]
        n = np.tanh(xin[t] @ p[G + "in/kernel"] + p[G + "in/bias"] + r * hn)
        h = (1 - zg) * n + zg * h0
        steps.append((h0, r, zg, hn, n))
        ys.append(h)
    Y = np.stack(ys)                                         # [T,B,H]
    q = Y @ p[f"Dense_{L}/kernel"] + p[f"Dense_{L}/bias"]    # :90
    if want_cache:
        return h, q, (trunk, xin, steps, Y)
    return h, q


def compute_targets(last_q, q_vals, reward, done, gamma, lam):
    """``_compute_targets`` (:295-323): Q(lambda) over the minibatch window.  q_vals/reward/done are the first T-1
    steps, last_q = max_a q[T-1]."""
    dt = q_vals.dtype
    g, l = dt.type(gamma), dt.type(lam)
    d = done.astype(dt)
    lam_ret = reward[-1] + g * (1 - d[-1]) * last_q
    next_q = q_vals[-1].max(-1)
    out = [None] * (reward.shape[0] - 1)
    for t in range(reward.shape[0] - 2, -1, -1):
        boot = reward[t] + g * (1 - d[t]) * next_q
        lam_ret_t = boot + g * l * (lam_ret - next_q)
        lam_ret_t = (1 - d[t]) * lam_ret_t + d[t] * reward[t]
        next_q = q_vals[t].max(-1)
        lam_ret = lam_ret_t
        out[t] = lam_ret_t
    # the scan emits the updated carry for t = T-3 .. 0; the initial lambda return (step T-2) is appended (:322)
    first = reward[-1] + g * (1 - d[-1]) * last_q
    return np.stack(out + [first]) if out else first[None]


def rnn_loss_and_grads(p, hs, obs, last_done, last_action, action, reward, done, gamma, lam):
    """``_loss_fn`` + ``value_and_grad`` (:330-366) on one minibatch window [T,B,...].
    -> loss, chosen_action_qvals [(T-1)*B], grads."""
    new_h, q, (trunk, xin, steps, Y) = rnn_forward(p, hs, obs, last_done, last_action, want_cache=True)
    dt = q.dtype
    T, B, A = q.shape
    L = _layers(p)
    last_q = q[-1].max(-1)                                                     # stop_gradient (:341-342)
    target = compute_targets(last_q, q[:-1], reward[:-1].astype(dt), done[:-1], gamma, lam).reshape(-1)
    qsa = np.take_along_axis(q, np.asarray(action, np.int64)[..., None], axis=-1)[..., 0]
    chosen = qsa[:-1].reshape(-1)
    diff = chosen - target
    loss = dt.type(0.5) * np.mean(diff * diff, dtype=dt)
    dq = np.zeros_like(q)
    np.put_along_axis(dq[:-1], np.asarray(action[:-1], np.int64)[..., None],
                      (diff / dt.type(diff.size)).reshape(T - 1, B, 1), axis=-1)
    g = {k: np.zeros_like(v) for k, v in p.items()}
    H = Y.shape[-1]
    g[f"Dense_{L}/kernel"] = Y.reshape(-1, H).T @ dq.reshape(-1, A)
    g[f"Dense_{L}/bias"] = dq.reshape(-1, A).sum(0)
    dY = dq @ p[f"Dense_{L}/kernel"].T
    dxin = np.zeros_like(xin)
    dh = np.zeros((B, H), dt)                                                  # gradient w.r.t. the carry
    for t in range(T - 1, -1, -1):                                             # BPTT through the scanned GRU
        h0, r, zg, hn, n = steps[t]
        dh = dh + dY[t]
        dn = dh * (1 - zg)
        dzg = dh * (h0 - n)
        dh0 = dh * zg
        da_n = dn * (1 - n * n)
        dr = da_n * hn
        dhn = da_n * r
        da_r = dr * r * (1 - r)
        da_z = dzg * zg * (1 - zg)
        for gate, da in (("ir", da_r), ("iz", da_z), ("in", da_n)):
            g[G + gate + "/kernel"] += xin[t].T @ da
            g[G + gate + "/bias"] += da.sum(0)
            dxin[t] += da @ p[G + gate + "/kernel"].T
        g[G + "hr/kernel"] += h0.T @ da_r
        g[G + "hz/kernel"] += h0.T @ da_z
        g[G + "hn/kernel"] += h0.T @ dhn
        g[G + "hn/bias"] += dhn.sum(0)
        dh0 = dh0 + da_r @ p[G + "hr/kernel"].T + da_z @ p[G + "hz/kernel"].T + dhn @ p[G + "hn/kernel"].T
        dh = np.where(np.asarray(last_done[t], bool)[:, None], dt.type(0), dh0)  # reset cuts the carry
    dx = dxin[..., :H]                                                         # the one-hot part has no parameters
    for l in reversed(range(L)):
        x_in, c, y = trunk[l]
        dy = dx * (y > 0)
        dz, g[f"LayerNorm_{l}/scale"], g[f"LayerNorm_{l}/bias"] = _layer_norm_bwd(dy, c, p[f"LayerNorm_{l}/scale"])
        F = x_in.shape[-1]
        g[f"Dense_{l}/kernel"] = x_in.reshape(-1, F).T @ dz.reshape(-1, H)
        g[f"Dense_{l}/bias"] = dz.reshape(-1, H).sum(0)
        dx = dz @ p[f"Dense_{l}/kernel"].T
    return loss, chosen, g
