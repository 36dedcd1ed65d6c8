"""CPU tests of oracle/pqn_rnn_ref.py (GRU PQN network + in-loss Q(lambda), purejaxql/pqn_rnn_gymnax.py): BPTT
agrees with central finite differences in fp64, the scanned GRU equals a step-by-step application with resets, and
the targets equal a literal transcription of the reference's reverse scan."""
import numpy as np

from oracle import pqn_ref as R
from oracle import pqn_rnn_ref as N

F64 = np.float64


def _setup(rng, T=5, B=4, D=3, A=2, H=8, layers=2):
    p = R.random_params(N.rnn_param_shapes(D, A, H, layers), seed=3, dtype=F64)
    hs = rng.standard_normal((B, H)) * 0.5
    obs = rng.standard_normal((T, B, D))
    last_done = rng.random((T, B)) < 0.25
    last_action = rng.integers(0, A, (T, B))
    action = rng.integers(0, A, (T, B))
    reward = rng.standard_normal((T, B))
    done = rng.random((T, B)) < 0.25
    return p, hs, obs, last_done, last_action, action, reward, done


def test_rnn_grads_match_finite_differences():
    rng = np.random.default_rng(0)
    p, hs, obs, ld, la, act, rew, done = _setup(rng)
    loss0, chosen0, g = N.rnn_loss_and_grads(p, hs, obs, ld, la, act, rew, done, 0.99, 0.65)
    # the targets are stop_gradient'ed (:341): freeze them at the base point for the finite differences
    q0 = N.rnn_forward(p, hs, obs, ld, la)[1]
    target = N.compute_targets(q0[-1].max(-1), q0[:-1], rew[:-1], done[:-1], 0.99, 0.65).reshape(-1)

    def frozen_loss(pp):
        q = N.rnn_forward(pp, hs, obs, ld, la)[1]
        chosen = np.take_along_axis(q, act[..., None], axis=-1)[..., 0][:-1].reshape(-1)
        return 0.5 * np.mean((chosen - target) ** 2)

    assert np.isclose(frozen_loss(p), loss0, rtol=0, atol=1e-15)
    h = 1e-6
    for k in p:
        if k.startswith("BatchNorm_0"):
            assert not g[k].any()      # dummy input BatchNorm (:74-76)
            continue
        flat = p[k].reshape(-1)
        for idx in rng.choice(flat.size, size=min(5, flat.size), replace=False):
            old = flat[idx]
            flat[idx] = old + h
            lp = frozen_loss(p)
            flat[idx] = old - h
            lm = frozen_loss(p)
            flat[idx] = old
            fd, an = (lp - lm) / (2 * h), g[k].reshape(-1)[idx]
            assert abs(fd - an) <= 3e-6 * max(1.0, abs(fd), abs(an)) + 1e-9, (k, idx, fd, an)


def test_scanned_gru_equals_stepwise_application_with_resets():
    rng = np.random.default_rng(1)
    p, hs, obs, ld, la, *_ = _setup(rng, T=6)
    h_all, q_all = N.rnn_forward(p, hs, obs, ld, la)
    h = hs
    for t in range(obs.shape[0]):  # the rollout applies the network one (dummy-time) step at a time (:229-241)
        h, q = N.rnn_forward(p, h, obs[t:t + 1], ld[t:t + 1], la[t:t + 1])
        assert np.allclose(q[0], q_all[t], rtol=0, atol=1e-12)
    assert np.allclose(h, h_all, rtol=0, atol=1e-12)
    # a reset makes the step independent of the incoming carry
    ld2 = ld.copy(); ld2[0] = True
    a = N.rnn_forward(p, hs, obs, ld2, la)[1]
    b = N.rnn_forward(p, hs + 7.0, obs, ld2, la)[1]
    assert np.array_equal(a, b)


def test_targets_equal_reference_reverse_scan_transcription():
    rng = np.random.default_rng(2)
    T, B, A, gamma, lam = 7, 5, 3, 0.99, 0.65
    q = rng.standard_normal((T, B, A)); reward = rng.standard_normal((T, B)); done = (rng.random((T, B)) < 0.3)
    last_q = q[-1].max(-1)
    got = N.compute_targets(last_q, q[:-1], reward[:-1], done[:-1], gamma, lam)
    # literal transcription of :296-323 (reverse lax.scan over all but the last element, then concatenate)
    rq, rr, rd = q[:-1], reward[:-1], done[:-1].astype(F64)
    lambda_returns = rr[-1] + gamma * (1 - rd[-1]) * last_q
    carry = (lambda_returns, rq[-1].max(-1))
    outs = []
    for t in range(rr.shape[0] - 2, -1, -1):
        lr_, nq = carry
        tb = rr[t] + gamma * (1 - rd[t]) * nq
        lr_ = tb + gamma * lam * (lr_ - nq)
        lr_ = (1 - rd[t]) * lr_ + rd[t] * rr[t]
        carry = (lr_, rq[t].max(-1))
        outs.append(lr_)
    want = np.concatenate([np.stack(outs[::-1]), lambda_returns[None]])
    assert got.shape == (T - 1, B) and np.allclose(got, want, rtol=0, atol=1e-13)


def test_gru_step_matches_torch_grucell():
    """Independent pin of the cell arithmetic: torch.nn.GRUCell uses the same gate equations (with two extra recurrent
    biases, set to zero here); weights are mapped from the flax layout [in, out] to torch's [3H, in] (r, z, n)."""
    import torch
    rng = np.random.default_rng(5)
    D, A, H, B = 8, 2, 8, 6                                                  # no trunk (layers=0): obs (width H) feeds the GRU
    p = R.random_params(N.rnn_param_shapes(D, A, H, 0), seed=9, dtype=F64)
    G = N.G
    cell = torch.nn.GRUCell(D + A, H).double()
    with torch.no_grad():
        cell.weight_ih.copy_(torch.tensor(np.concatenate([p[G + g + "/kernel"].T for g in ("ir", "iz", "in")])))
        cell.weight_hh.copy_(torch.tensor(np.concatenate([p[G + g + "/kernel"].T for g in ("hr", "hz", "hn")])))
        cell.bias_ih.copy_(torch.tensor(np.concatenate([p[G + g + "/bias"] for g in ("ir", "iz", "in")])))
        cell.bias_hh.copy_(torch.tensor(np.concatenate([np.zeros(H), np.zeros(H), p[G + "hn/bias"]])))
    hs = rng.standard_normal((B, H))
    obs = rng.standard_normal((1, B, H))          # with layers=0 the "trunk output" is the observation itself
    la = rng.integers(0, A, (1, B))
    p0 = dict(p)
    # layers=0 => Dense_0 is the head; rnn_forward reads its input width from the parameter shapes
    new_h, q = N.rnn_forward(p0, hs, obs[..., :D], np.zeros((1, B), bool), la)
    onehot = np.eye(A)[la[0]]
    want = cell(torch.tensor(np.concatenate([obs[0, :, :D], onehot], -1)), torch.tensor(hs)).detach().numpy()
    assert np.allclose(new_h, want, rtol=0, atol=1e-12)
    assert np.allclose(q[0], want @ p["Dense_0/kernel"] + p["Dense_0/bias"], rtol=0, atol=1e-12)
