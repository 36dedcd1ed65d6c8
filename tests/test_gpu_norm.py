"""GPU parity of the NORM_TYPE / NORM_INPUT network variants (purejaxql/pqn_minatar.py:24-69,
pqn_gymnax.py:29-58) against oracle/pqn_ref_norm.py: eval-mode forward (running statistics) to 1e-5, training
loss/gradients to 2e-5 of the gradient's scale against the fp64 oracle, the batch_stats side effects, and one whole
update through make_train/train."""
import numpy as np
import pytest
import torch

from oracle import gymnax_envs as G
from oracle import jax_prng as jr
from oracle import pqn_ref as R
from oracle import pqn_ref_norm as RN

pytestmark = pytest.mark.gpu

# biases that feed a BatchNorm directly (their exact gradient is zero)
BN_DEAD_BIASES = {"cnn": ("CNN_0/Conv_0/bias", "CNN_0/Dense_0/bias"), "mlp": ("Dense_0/bias", "Dense_1/bias")}

VARIANTS = [("batch_norm", False), ("batch_norm", True), ("none", False), ("none", True), ("layer_norm", True)]


def dev():
    return torch.device("cuda:0")


def _pack(obs_bool):
    n = obs_bool.shape[0]
    flat = obs_bool.reshape(n, -1).astype(np.uint8)
    nb = flat.shape[1]
    pw = ((nb + 31) // 32 + 3) // 4 * 4
    padded = np.zeros((n, pw * 32), np.uint8)
    padded[:, :nb] = flat
    return np.ascontiguousarray(np.packbits(padded, axis=-1, bitorder="little")).view("<u4").view(np.int32)


def _minatar_obs(name, n, seed):
    env = G.make(name, log=False)
    obs, st = env.reset(jr.split(jr.PRNGKey(seed), n))
    rng = np.random.default_rng(seed)
    for t in range(25):
        obs, st, *_ = env.step(jr.split(jr.PRNGKey(seed * 1000 + t), n), st,
                               rng.integers(0, env.num_actions, n).astype(np.int32))
    return obs


def _rand_stats(stats, seed):
    rng = np.random.default_rng(seed)
    return {k: {"mean": (0.1 * rng.standard_normal(v["mean"].shape)).astype(np.float32),
                "var": (0.5 + rng.random(v["var"].shape)).astype(np.float32)} for k, v in stats.items()}


def _setup(kind, norm_type, norm_input, S):
    from purejaxql_b200.networks import NET_CNN, NET_MLP, QNetworkSpec
    if kind == "cnn":
        C, A = 4, 3
        spec = QNetworkSpec(NET_CNN, C, A, norm_type=norm_type, norm_input=norm_input)
        shapes = RN.cnn_param_shapes(C, A, norm_type)
        stats0 = RN.cnn_batch_stats(C, norm_type)
    else:
        D, A, H, Ls = 4, 2, 128, 2
        spec = QNetworkSpec(NET_MLP, D, A, H, Ls, norm_type=norm_type, norm_input=norm_input)
        shapes = RN.mlp_param_shapes(D, A, H, Ls, norm_type)
        stats0 = RN.mlp_batch_stats(D, H, Ls, norm_type)
    assert set("/".join(p) for p, *_ in spec.entries) == set(shapes), (sorted(shapes), spec.flat_names("/"))
    ps = [R.random_params(shapes, 30 + s) for s in range(S)]
    if norm_type == "batch_norm":
        # a bias in front of a BatchNorm is a no-op; a non-zero one only makes flax's fast variance E[x^2] - E[x]^2
        # cancel catastrophically in fp32 (z = 0.1 +- 1e-3 with NORM_INPUT=False), which the fp64 oracle does not share
        for p in ps:
            for k in p:
                if k.endswith("/bias") and ("Conv_0" in k or (k.startswith("CNN_0/Dense_0") or k in ("Dense_0/bias", "Dense_1/bias"))) \
                        and not k.startswith("Dense_%d" % (2 if kind == "mlp" else 0)):
                    p[k] = np.zeros_like(p[k])
    sts = [_rand_stats(stats0, 50 + s) for s in range(S)]
    flat = torch.cat([spec.flatten(p, 1, dev()) for p in ps], 0).contiguous()
    stf = torch.cat([spec.flatten_stats(st, 1, dev()) for st in sts], 0).contiguous()
    return spec, ps, sts, flat, stf


def _ws(spec, S, rows):
    from purejaxql_b200 import _lib
    return torch.empty(int(_lib.lib().pqn_net_workspace_bytes(spec.desc, S, rows)), dtype=torch.uint8, device=dev())


def _inputs(kind, S, rows):
    if kind == "cnn":
        obs = np.stack([_minatar_obs("Breakout-MinAtar", rows, s + 1) for s in range(S)])
        dev_obs = torch.from_numpy(np.stack([_pack(obs[s] != 0) for s in range(S)])).to(dev()).contiguous()
    else:
        obs = np.random.default_rng(4).standard_normal((S, rows, 4)).astype(np.float32) * np.array([1, 2, .2, 3], np.float32)
        dev_obs = torch.from_numpy(obs).to(dev()).contiguous()
    return obs, dev_obs


@pytest.mark.parametrize("kind", ["cnn", "mlp"])
@pytest.mark.parametrize("norm_type,norm_input", VARIANTS)
def test_norm_variant_eval_forward_matches_oracle(kind, norm_type, norm_input):
    from purejaxql_b200 import _lib
    S, rows = 2, 203
    spec, ps, sts, flat, stf = _setup(kind, norm_type, norm_input, S)
    obs, dev_obs = _inputs(kind, S, rows)
    A = spec.num_actions
    q = torch.zeros((S * rows, A), device=dev())
    ws = _ws(spec, S, rows)
    _lib.check(_lib.lib().pqn_qnet_forward(spec.desc, _lib.p(flat), _lib.p(stf), _lib.p(dev_obs), None, rows, _lib.p(q),
                                           S, rows, _lib.p(ws), _lib.stream_ptr()), "pqn_qnet_forward")
    torch.cuda.synchronize()
    q = q.cpu().numpy().reshape(S, rows, A)
    fwd = RN.cnn_forward if kind == "cnn" else RN.mlp_forward
    for s in range(S):
        ref, _ = fwd(ps[s], sts[s], obs[s].astype(np.float32), False, norm_type, norm_input)
        assert np.abs(q[s] - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), (s, np.abs(q[s] - ref).max())


@pytest.mark.parametrize("kind", ["cnn", "mlp"])
@pytest.mark.parametrize("norm_type,norm_input", VARIANTS)
def test_norm_variant_loss_grad_matches_fp64_oracle(kind, norm_type, norm_input):
    from purejaxql_b200 import _lib
    S, total, rows = 2, 300, 256
    spec, ps, sts, flat, stf = _setup(kind, norm_type, norm_input, S)
    obs, dev_obs = _inputs(kind, S, total)
    rng = np.random.default_rng(7)
    A, F = spec.num_actions, spec.in_c
    gather = np.stack([rng.permutation(total)[:rows] for _ in range(S)]).astype(np.int32)
    act = rng.integers(0, A, (S, total)).astype(np.int32)
    tgt = rng.standard_normal((S, total)).astype(np.float32)
    grads = torch.zeros_like(flat)
    ls = torch.zeros(S, device=dev()); qs = torch.zeros(S, device=dev())
    bn = torch.zeros((S, 2 * F), device=dev())
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev(), dt)
    st_dev = stf.clone()
    L = _lib.lib()
    # keep every device buffer referenced until the call has run (a temporary would be recycled by the allocator)
    tg_, ta_, tt_, ws = t(gather, torch.int32), t(act, torch.int32), t(tgt, torch.float32), _ws(spec, S, rows)
    _lib.check(L.pqn_qnet_loss_grad(spec.desc, _lib.p(flat), _lib.p(st_dev), _lib.p(dev_obs), _lib.p(tg_),
                                    total, _lib.p(ta_), _lib.p(tt_), total,
                                    _lib.p(grads), _lib.p(ls), _lib.p(qs), _lib.p(bn), S, rows, _lib.p(ws),
                                    _lib.stream_ptr()), "pqn_qnet_loss_grad")
    count = float(rows * (100 if kind == "cnn" else 1))
    _lib.check(L.pqn_bn_stats_update(_lib.p(st_dev), _lib.p(bn), S, F, spec.stats_total, count, 0.99, _lib.stream_ptr()))
    torch.cuda.synchronize()
    gtree = spec.unflatten(grads)
    sttree = spec.unflatten_stats(st_dev)
    lg = RN.cnn_loss_and_grads if kind == "cnn" else RN.mlp_loss_and_grads
    for s in range(S):
        p64 = {k: v.astype(np.float64) for k, v in ps[s].items()}
        st64 = {k: {kk: vv.astype(np.float64) for kk, vv in v.items()} for k, v in sts[s].items()}
        x = obs[s][gather[s]].astype(np.float64)
        loss, q_sa, g, new_stats = lg(p64, st64, x, act[s][gather[s]], tgt[s][gather[s]].astype(np.float64), norm_type,
                                      norm_input)
        # fp32 batch statistics (E[x^2] - E[x]^2 over 16k elements) against the fp64 oracle: a few 1e-5 on q
        assert abs(float(ls[s]) - loss) < 5e-5 * max(1.0, abs(loss)), (float(ls[s]), loss)
        assert abs(float(qs[s]) - q_sa.mean()) < 5e-5 * max(1.0, abs(q_sa.mean()))
        scale = max(np.abs(v).max() for v in g.values())
        errs = {}
        for path, *_ in spec.entries:
            d = gtree
            for k in path:
                d = d[k]
            got, want = d[s].cpu().numpy(), g["/".join(path)]
            name = "/".join(path)
            tol = 2e-5
            if norm_type == "batch_norm":
                # fp32 batch statistics: the NumPy oracle run in fp32 already differs from its fp64 self by up to 6e-5
                # of the scale here (x/255 inputs, var ~ eps), and by 1e-2 on a bias in front of a BatchNorm, whose
                # exact gradient is zero (the sum of a batch-normalised gradient)
                tol = 5e-2 if name in BN_DEAD_BIASES[kind] else 2e-4
            errs[name] = (float(np.abs(got - want).max() / scale), tol)
        bad = {k: v for k, v in errs.items() if not v[0] < v[1]}
        assert not bad, (bad, errs)
        for path, off, n in spec.stats_entries():
            d = sttree
            for k in path:
                d = d[k]
            want = new_stats["/".join(path)]
            assert np.allclose(d["mean"][s].cpu().numpy(), want["mean"], atol=2e-6), path
            assert np.allclose(d["var"][s].cpu().numpy(), want["var"], atol=2e-6), path


@pytest.mark.parametrize("env_name,kind,module,flatten,norm_type,norm_input", [
    ("Breakout-MinAtar", "cnn", "pqn_minatar", False, "batch_norm", False),
    ("Breakout-MinAtar", "cnn", "pqn_minatar", False, "layer_norm", True),
    ("Breakout-MinAtar", "cnn", "pqn_minatar", False, "none", False),
    ("CartPole-v1", "mlp", "pqn_gymnax", True, "batch_norm", True),
])
def test_norm_variant_update_step_matches_oracle(env_name, kind, module, flatten, norm_type, norm_input, monkeypatch):
    """One whole `_update_step` through make_train/train with the variant network against the oracle's update step
    (eps = 1: integer-exact rollouts).  The oracle's update_step is reused with stateful forward / loss closures that
    carry the batch_stats collection (train=False in the rollout, mutable batch_stats in the loss, :277-296)."""
    import importlib
    mod = importlib.import_module(f"purejaxql_b200.{module}")
    cfg = dict(ENV_NAME=env_name, NUM_ENVS=64, NUM_STEPS=8, NUM_MINIBATCHES=4, NUM_EPOCHS=2, EPS_START=1.0, EPS_FINISH=1.0,
               EPS_DECAY=0.1, LR=5e-4, MAX_GRAD_NORM=10, GAMMA=0.99, LAMBDA=0.65, NORM_TYPE=norm_type,
               NORM_INPUT=norm_input, LR_LINEAR_DECAY=True, WANDB_MODE="disabled", TEST_DURING_TRAINING=False,
               HIDDEN_SIZE=128, NUM_LAYERS=2, CUDA_GRAPH=False)
    cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(2 * cfg["NUM_STEPS"] * cfg["NUM_ENVS"])
    train = mod.make_train(cfg)
    eng = train.engine
    S = 2
    rngs = jr.split(jr.PRNGKey(21), S)
    cap = {}
    orig = eng.spec.init
    eng.spec.init = lambda k, d: cap.setdefault("flat", orig(k, d)).clone()
    out = train(rngs)
    ts = out["runner_state"][0]
    tree0 = eng.spec.unflatten(cap["flat"])
    E = cfg["NUM_ENVS"]
    fwd0 = RN.cnn_forward if kind == "cnn" else RN.mlp_forward
    lg0 = RN.cnn_loss_and_grads if kind == "cnn" else RN.mlp_loss_and_grads
    for s in range(S):
        def leaf(tree, path):
            d = tree
            for k in path:
                d = d[k]
            return d[s].cpu().numpy()
        params = {"/".join(p): leaf(tree0, p).astype(np.float32) for p, *_ in eng.spec.entries}
        box = {"stats": (RN.cnn_batch_stats(eng.spec.in_c, norm_type) if kind == "cnn" else
                         RN.mlp_batch_stats(eng.spec.in_c, 128, 2, norm_type))}

        def fwd(p, obs, want_cache=False):
            return fwd0(p, box["stats"], obs, False, norm_type, norm_input)[0]

        def lossgrad(p, obs, a, t):
            loss, q_sa, g, box["stats"] = lg0(p, box["stats"], obs.astype(np.float32), a, t, norm_type, norm_input)
            return loss, q_sa, g
        monkeypatch.setattr(R, "cnn_forward" if kind == "cnn" else "mlp_forward", fwd)
        monkeypatch.setattr(R, "cnn_loss_and_grads" if kind == "cnn" else "mlp_loss_and_grads", lossgrad)
        K1 = jr.split(rngs[s], 2)[0]
        K2 = jr.split(K1, 2)[0]
        k = jr.split(K2, 2); K3, kR = k[0], k[1]
        env = G.make(env_name, flatten=flatten)
        obs, st = env.reset(jr.split(kR, E))
        rng = jr.split(K3, 2)[1]
        opt = R.opt_init(params)
        F = eng.spec.in_c
        bs = {"mean": np.zeros(F, np.float32), "var": np.ones(F, np.float32)}
        total = cfg["NUM_UPDATES_DECAY"] * cfg["NUM_MINIBATCHES"] * cfg["NUM_EPOCHS"]
        lr_fn = lambda i: R.linear_schedule(cfg["LR"], 1e-20, total, i)
        for u in range(2):
            params, opt, bs, obs, st, rng, m, tr, tg = R.update_step(env, kind, params, opt, bs, obs, st, rng, dict(cfg), u,
                                                                     lr_fn)
            tol = 2e-4 if kind == "cnn" else 2e-2
            got = float(out["metrics"]["td_loss"][s, u])
            assert abs(got - m["td_loss"]) < tol * max(1.0, abs(m["td_loss"])), (u, got, m["td_loss"])
        if kind == "cnn":
            for p, *_ in eng.spec.entries:
                if norm_type == "batch_norm" and p[-1] == "bias" and p[-2] in ("Conv_0", "Dense_0") and p[0] == "CNN_0":
                    continue   # a bias in front of a BatchNorm has an analytically zero gradient: RAdam normalises pure
                               # rounding noise there, so the two implementations random-walk apart by ~lr per step
                tol = 2e-4 if norm_type == "batch_norm" else 5e-5    # batch statistics in fp32 + RAdam on small gradients
                assert np.abs(leaf(ts.params, p) - params["/".join(p)]).max() < tol, p
            for path, off, n in eng.spec.stats_entries():
                want = box["stats"]["/".join(path)]
                d = ts.batch_stats
                for kk in path:
                    d = d[kk]
                # running statistics of a hidden BatchNorm see every parameter difference of the 2 x 8 optimiser steps (dot products over 1024 inputs)
                stol = 5e-4 if norm_type == "batch_norm" else 1e-5
                assert np.allclose(d["mean"][s].cpu().numpy(), want["mean"], atol=stol), path
                assert np.allclose(d["var"][s].cpu().numpy(), want["var"], atol=stol), path
            assert np.array_equal(out["runner_state"][3][s].cpu().numpy().view(np.uint32), rng)
