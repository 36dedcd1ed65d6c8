"""Self-consistency of the oracle's network / target / optimizer arithmetic."""
import numpy as np
import torch

from oracle import pqn_ref as R


def _fd_check(lossgrad, p, obs, act, tgt, names, rng):
    loss, _, g = lossgrad(p, obs, act, tgt)
    for name in names:
        for _ in range(3):
            idx = tuple(rng.integers(0, s) for s in p[name].shape)
            h = 1e-6
            pp = {k: v.copy() for k, v in p.items()}
            pp[name][idx] += h
            lp = lossgrad(pp, obs, act, tgt)[0]
            pp[name][idx] -= 2 * h
            lm = lossgrad(pp, obs, act, tgt)[0]
            fd = (lp - lm) / (2 * h)
            assert abs(fd - g[name][idx]) < 1e-6 + 1e-4 * abs(fd), (name, idx, fd, g[name][idx])


def test_cnn_grads_match_finite_differences_f64():
    rng = np.random.default_rng(0)
    p = R.random_params(R.cnn_param_shapes(4, 3), 1, np.float64)
    obs = (rng.random((6, 10, 10, 4)) < 0.2).astype(np.float64) * 255.0  # scale so /255 keeps O(1) signal
    act = rng.integers(0, 3, 6)
    tgt = rng.standard_normal(6)
    names = [k for k in p if not k.startswith("BatchNorm")]
    _fd_check(R.cnn_loss_and_grads, p, obs, act, tgt, names, rng)


def test_mlp_grads_match_finite_differences_f64():
    rng = np.random.default_rng(0)
    p = R.random_params(R.mlp_param_shapes(4, 2, 32, 2), 1, np.float64)
    obs = rng.standard_normal((8, 4))
    act = rng.integers(0, 2, 8)
    tgt = rng.standard_normal(8)
    names = [k for k in p if not k.startswith("BatchNorm")]
    _fd_check(R.mlp_loss_and_grads, p, obs, act, tgt, names, rng)


def test_cnn_forward_matches_torch_fp32():
    rng = np.random.default_rng(2)
    p = R.random_params(R.cnn_param_shapes(4, 3), 3)
    obs = (rng.random((16, 10, 10, 4)) < 0.15).astype(np.float32)
    q = R.cnn_forward(p, obs)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    x = t(obs).permute(0, 3, 1, 2) / 255.0
    w = t(p["CNN_0/Conv_0/kernel"]).permute(3, 2, 0, 1)
    z = torch.nn.functional.conv2d(x, w, t(p["CNN_0/Conv_0/bias"])).permute(0, 2, 3, 1)
    z = torch.nn.functional.layer_norm(z, (16,), t(p["CNN_0/LayerNorm_0/scale"]), t(p["CNN_0/LayerNorm_0/bias"]), 1e-6)
    h = torch.relu(z).reshape(16, -1)
    z = h @ t(p["CNN_0/Dense_0/kernel"]) + t(p["CNN_0/Dense_0/bias"])
    z = torch.nn.functional.layer_norm(z, (128,), t(p["CNN_0/LayerNorm_1/scale"]), t(p["CNN_0/LayerNorm_1/bias"]), 1e-6)
    qq = torch.relu(z) @ t(p["Dense_0/kernel"]) + t(p["Dense_0/bias"])
    assert np.abs(q - qq.numpy()).max() < 2e-5


def test_q_lambda_matches_naive_loop():
    rng = np.random.default_rng(5)
    T, N, A = 7, 9, 3
    r = rng.standard_normal((T, N)).astype(np.float32)
    d = rng.random((T, N)) < 0.3
    qv = rng.standard_normal((T, N, A)).astype(np.float32)
    lq = rng.standard_normal(N).astype(np.float32)
    out = R.q_lambda_targets(r, d, qv, lq, 0.99, 0.65)
    g, lam = 0.99, 0.65
    for n in range(N):
        nq = lq[n] * (1 - d[-1, n])
        G = r[-1, n] + g * nq
        assert abs(out[-1, n] - G) < 1e-5
        for t in range(T - 2, -1, -1):
            boot = r[t, n] + g * (1 - d[t, n]) * nq
            G = boot + g * lam * (G - nq)
            G = (1 - d[t, n]) * G + d[t, n] * r[t, n]
            nq = qv[t, n].max()
            assert abs(out[t, n] - G) < 1e-4


def test_radam_rectification_threshold_and_clip():
    assert not R.radam_scalars(5)[3] and R.radam_scalars(6)[3]
    p = {"w": np.ones(4, np.float32)}
    g = {"w": np.full(4, 100.0, np.float32)}
    newp, opt, gn = R.radam_clip_step(p, g, R.opt_init(p), 0.1, 10.0)
    assert abs(gn - 200.0) < 1e-3
    # step 1: un-rectified => update = m_hat = clipped grad (norm 10 => 5 per element)
    assert np.allclose(newp["w"], 1.0 - 0.1 * 5.0, atol=1e-6)


def test_linear_schedule():
    assert R.linear_schedule(1.0, 0.05, 7.6, 0) == np.float32(1.0)
    assert abs(R.linear_schedule(1.0, 0.05, 7.6, 100) - 0.05) < 1e-7
    assert abs(R.linear_schedule(1.0, 0.05, 10, 5) - 0.525) < 1e-6


def test_radam_matches_torch_radam_over_the_rectification_switch():
    """Independent pin of the optimizer arithmetic: torch.optim.RAdam implements the same algorithm as
    optax.scale_by_radam (they differ only in where eps enters the bias-corrected denominator, an O(eps) effect)."""
    import torch
    rng = np.random.default_rng(0)
    w0 = rng.standard_normal(16)
    p = {"w": w0.copy()}
    opt = R.opt_init(p)
    tw = torch.tensor(w0.copy(), dtype=torch.float64, requires_grad=True)
    topt = torch.optim.RAdam([tw], lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    for step in range(12):  # steps 1..5 un-rectified, 6.. rectified
        g = rng.standard_normal(16)
        p, opt, _ = R.radam_clip_step(p, {"w": g.copy()}, opt, 3e-3, 1e9)
        tw.grad = torch.tensor(g.copy(), dtype=torch.float64)
        topt.step()
        assert np.allclose(p["w"], tw.detach().numpy(), rtol=1e-6, atol=1e-9), step


def test_cnn_and_mlp_grads_match_torch_autograd_f64():
    """The oracle's hand-written backward against an independent autodiff (torch, fp64) of the same network built from
    torch primitives (conv2d / layer_norm / relu / matmul), for the MinAtar CNN and the gymnax MLP."""
    rng = np.random.default_rng(11)
    F64 = np.float64
    tt = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, requires_grad=True)
    # ---- CNN
    B, C, A = 9, 4, 3
    p = R.random_params(R.cnn_param_shapes(C, A), 5, dtype=F64)
    obs = (rng.random((B, 10, 10, C)) < 0.2).astype(F64)
    act, tgt = rng.integers(0, A, B), rng.standard_normal(B)
    loss, q_sa, g = R.cnn_loss_and_grads(p, obs, act, tgt)
    tp = {k: tt(v) for k, v in p.items()}
    x = torch.tensor(obs).permute(0, 3, 1, 2) / 255.0
    z = torch.nn.functional.conv2d(x, tp["CNN_0/Conv_0/kernel"].permute(3, 2, 0, 1), tp["CNN_0/Conv_0/bias"]).permute(0, 2, 3, 1)
    z = torch.nn.functional.layer_norm(z, (16,), tp["CNN_0/LayerNorm_0/scale"], tp["CNN_0/LayerNorm_0/bias"], 1e-6)
    h = torch.relu(z).reshape(B, -1)
    z = torch.nn.functional.layer_norm(h @ tp["CNN_0/Dense_0/kernel"] + tp["CNN_0/Dense_0/bias"], (128,),
                                       tp["CNN_0/LayerNorm_1/scale"], tp["CNN_0/LayerNorm_1/bias"], 1e-6)
    q = torch.relu(z) @ tp["Dense_0/kernel"] + tp["Dense_0/bias"]
    tl = 0.5 * ((q[torch.arange(B), torch.tensor(act)] - torch.tensor(tgt)) ** 2).mean()
    tl.backward()
    assert abs(float(tl.detach()) - loss) < 1e-12
    for k in p:
        ref = tp[k].grad.numpy() if tp[k].grad is not None else np.zeros_like(p[k])
        assert np.allclose(g[k], ref, rtol=1e-9, atol=1e-12), k
    # ---- MLP
    D, H = 6, 32
    p = R.random_params(R.mlp_param_shapes(D, A, H, 2), 6, dtype=F64)
    x = rng.standard_normal((B, D))
    loss, q_sa, g = R.mlp_loss_and_grads(p, x, act, tgt)
    tp = {k: tt(v) for k, v in p.items()}
    h = torch.tensor(x)
    for l in range(2):
        z = h @ tp[f"Dense_{l}/kernel"] + tp[f"Dense_{l}/bias"]
        h = torch.relu(torch.nn.functional.layer_norm(z, (H,), tp[f"LayerNorm_{l}/scale"], tp[f"LayerNorm_{l}/bias"], 1e-6))
    q = h @ tp["Dense_2/kernel"] + tp["Dense_2/bias"]
    tl = 0.5 * ((q[torch.arange(B), torch.tensor(act)] - torch.tensor(tgt)) ** 2).mean()
    tl.backward()
    assert abs(float(tl.detach()) - loss) < 1e-12
    for k in p:
        ref = tp[k].grad.numpy() if tp[k].grad is not None else np.zeros_like(p[k])
        assert np.allclose(g[k], ref, rtol=1e-9, atol=1e-12), k
