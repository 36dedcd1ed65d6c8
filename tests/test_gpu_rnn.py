"""GPU parity of the recurrent (GRU) PQN path (purejaxql/pqn_rnn_gymnax.py) against
oracle/pqn_rnn_ref.py: one network step (rollout form), window loss + BPTT gradients vs the fp64 oracle, and whole
updates through make_train/train (memory warm-up, env-axis minibatches, in-loss Q(lambda)) vs an oracle replay."""
import numpy as np
import pytest
import torch

from oracle import gymnax_envs as G
from oracle import jax_prng as jr
from oracle import pqn_ref as R
from oracle import pqn_rnn_ref as RR

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def t_(a, dt):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev(), dt)


def _setup(S, D=4, A=2, H=128, Ls=2):
    from purejaxql_b200.networks import NET_RNN, QNetworkSpec
    spec = QNetworkSpec(NET_RNN, D, A, H, Ls)
    shapes = RR.rnn_param_shapes(D, A, H, Ls)
    ps = [R.random_params(shapes, 70 + s) for s in range(S)]
    for p in ps:                                   # recurrent kernels at a realistic (orthogonal-like) scale
        for g in ("hr", "hz", "hn"):
            p[RR.G + g + "/kernel"] = (p[RR.G + g + "/kernel"] * 0.5).astype(np.float32)
    flat = torch.cat([spec.flatten(p, 1, dev()) for p in ps], 0).contiguous()
    return spec, ps, flat


def _ws(spec, S, rows):
    from purejaxql_b200 import _lib
    return torch.empty(int(_lib.lib().pqn_net_workspace_bytes(spec.desc, S, rows)), dtype=torch.uint8, device=dev())


@pytest.mark.parametrize("H,Ls", [(128, 2), (256, 2), (128, 1)])
def test_rnn_step_matches_oracle(H, Ls):
    from purejaxql_b200 import _lib
    S, E, D, A = 2, 37, 4, 2
    spec, ps, flat = _setup(S, D, A, H, Ls)
    rng = np.random.default_rng(0)
    hs = rng.standard_normal((S, E, H)).astype(np.float32) * 0.5
    obs = rng.standard_normal((S, E, D)).astype(np.float32)
    ld = (rng.random((S, E)) < 0.3)
    la = rng.integers(0, A, (S, E)).astype(np.int32)
    hs_d, obs_d, ld_d, la_d = t_(hs, torch.float32), t_(obs, torch.float32), t_(ld.astype(np.uint8), torch.uint8), t_(la, torch.int32)
    q = torch.zeros((S * E, A), device=dev())
    ws = _ws(spec, S, E)
    _lib.check(_lib.lib().pqn_rnn_step(spec.desc, _lib.p(flat), _lib.p(hs_d), _lib.p(obs_d), E, _lib.p(ld_d), _lib.p(la_d),
                                       _lib.p(q), S, E, _lib.p(ws), _lib.stream_ptr()), "pqn_rnn_step")
    torch.cuda.synchronize()
    for s in range(S):
        new_h, qq = RR.rnn_forward(ps[s], hs[s], obs[s][None], ld[s][None], la[s][None])
        assert np.abs(q.cpu().numpy().reshape(S, E, A)[s] - qq[0]).max() < 1e-5
        assert np.abs(hs_d.cpu().numpy()[s] - new_h).max() < 1e-5


@pytest.mark.parametrize("H,Ls,T,B", [(128, 2, 12, 3), (256, 2, 9, 5), (128, 1, 6, 1)])
def test_rnn_loss_grad_matches_fp64_oracle(H, Ls, T, B):
    from purejaxql_b200 import _lib
    S, D, A = 2, 4, 2
    spec, ps, flat = _setup(S, D, A, H, Ls)
    rng = np.random.default_rng(1)
    hs0 = rng.standard_normal((S, B, H)).astype(np.float32) * 0.5
    obs = rng.standard_normal((S, T, B, D)).astype(np.float32)
    ld = rng.random((S, T, B)) < 0.15
    la = rng.integers(0, A, (S, T, B)).astype(np.int32)
    ac = rng.integers(0, A, (S, T, B)).astype(np.int32)
    rw = (rng.random((S, T, B)) * 0.1).astype(np.float32)
    dn = rng.random((S, T, B)) < 0.15
    bufs = [t_(hs0, torch.float32), t_(obs, torch.float32), t_(ld.astype(np.uint8), torch.uint8), t_(la, torch.int32),
            t_(ac, torch.int32), t_(rw, torch.float32), t_(dn.astype(np.uint8), torch.uint8)]
    grads = torch.zeros_like(flat)
    ls, qs = torch.zeros(S, device=dev()), torch.zeros(S, device=dev())
    ws = _ws(spec, S, T * B)
    _lib.check(_lib.lib().pqn_rnn_loss_grad(spec.desc, _lib.p(flat), *[_lib.p(b) for b in bufs], _lib.p(grads), _lib.p(ls),
                                            _lib.p(qs), S, T, B, 0.99, 0.95, _lib.p(ws), _lib.stream_ptr()),
               "pqn_rnn_loss_grad")
    torch.cuda.synchronize()
    gtree = spec.unflatten(grads)
    for s in range(S):
        p64 = {k: v.astype(np.float64) for k, v in ps[s].items()}
        loss, chosen, g = RR.rnn_loss_and_grads(p64, hs0[s].astype(np.float64), obs[s].astype(np.float64), ld[s], la[s],
                                                ac[s], rw[s].astype(np.float64), dn[s], 0.99, 0.95)
        assert abs(float(ls[s]) - loss) < 1e-5 * max(1.0, abs(loss)), (float(ls[s]), loss)
        assert abs(float(qs[s]) - chosen.mean()) < 1e-5 * max(1.0, abs(chosen.mean()))
        scale = max(np.abs(v).max() for v in g.values())
        for path, *_ in spec.entries:
            d = gtree
            for k in path:
                d = d[k]
            got, want = d[s].cpu().numpy(), g["/".join(path)]
            assert np.abs(got - want).max() < 2e-5 * scale, (path, np.abs(got - want).max(), scale)


def _oracle_step(env, p, hs, obs, ld, la, st, rng, eps, rew_scale, E):
    """_step_env / _random_step (:192-236, :514-529) for one seed."""
    ks = jr.split(rng, 3)
    rng, rng_a, rng_s = ks[0], ks[1], ks[2]
    new_hs, q = RR.rnn_forward(p, hs, obs[None], ld[None], la[None])
    q = q[0]
    act = R.eps_greedy(jr.split(rng_a, E), q, eps)
    new_obs, st, reward, done, info = env.step(jr.split(rng_s, E), st, act)
    tr = dict(last_hs=hs, obs=obs, action=act, reward=(np.float32(rew_scale) * reward).astype(np.float32), done=done,
              last_done=ld, last_action=la)
    return (new_hs.astype(np.float32), new_obs, done, act, st, rng), tr, info


def test_rnn_update_steps_match_oracle():
    """Two whole updates of pqn_rnn_gymnax.make_train/train with eps = 1 (random actions => the rollouts do not depend
    on the network; CartPole physics fp32 both sides) against the oracle: losses per update, final parameters."""
    from purejaxql_b200 import pqn_rnn_gymnax
    cfg = dict(ENV_NAME="CartPole-v1", NUM_ENVS=8, NUM_STEPS=12, MEMORY_WINDOW=3, NUM_MINIBATCHES=4, NUM_EPOCHS=2,
               EPS_START=1.0, EPS_FINISH=1.0, EPS_DECAY=0.2, LR=1e-4, MAX_GRAD_NORM=10, GAMMA=0.99, LAMBDA=0.95,
               NORM_TYPE="layer_norm", NORM_INPUT=False, HIDDEN_SIZE=128, NUM_LAYERS=2, LR_LINEAR_DECAY=True, REW_SCALE=0.1,
               WANDB_MODE="disabled", TEST_DURING_TRAINING=False)
    nupd = 2
    cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(nupd * cfg["NUM_STEPS"] * cfg["NUM_ENVS"])
    train = pqn_rnn_gymnax.make_train(cfg)
    eng = train.engine
    S = 2
    rngs = jr.split(jr.PRNGKey(31), S)
    cap = {}
    orig = eng.spec.init
    eng.spec.init = lambda k, d: cap.setdefault("flat", orig(k, d)).clone()
    out = train(rngs)
    ts = out["runner_state"][0]
    tree0 = eng.spec.unflatten(cap["flat"])
    T, E, W, nmb = cfg["NUM_STEPS"], cfg["NUM_ENVS"], cfg["MEMORY_WINDOW"], cfg["NUM_MINIBATCHES"]
    Bm = E // nmb
    H = 128
    for s in range(S):
        def leaf(tree, path):
            d = tree
            for k in path:
                d = d[k]
            return d[s].cpu().numpy()
        params = {"/".join(p): leaf(tree0, p).astype(np.float32) for p, *_ in eng.spec.entries}
        env = G.make("CartPole-v1", flatten=True)
        k = jr.split(rngs[s], 2); rng = k[0]                               # :255 (init key = rng)
        k = jr.split(rng, 2); rng = k[0]                                   # :505 (test key unused)
        k = jr.split(rng, 2); rng, kR = k[0], k[1]                         # :508
        obs, st = env.reset(jr.split(kR, E))
        hs = np.zeros((E, H), np.float32); ld = np.zeros(E, bool); la = np.zeros(E, np.int32)
        k = jr.split(rng, 2); carry = k[1]                                 # :531
        mem = []
        for _ in range(W + T):
            (hs, obs, ld, la, st, carry), tr, _ = _oracle_step(env, params, hs, obs, ld, la, st, carry, 1.0, 0.1, E)
            mem.append(tr)
        rng = carry                                                        # re-binding (:532)
        k = jr.split(rng, 2); rng = k[1]                                   # :541
        opt = R.opt_init(params)
        total = cfg["NUM_UPDATES_DECAY"] * nmb * cfg["NUM_EPOCHS"]
        lr_fn = lambda i: R.linear_schedule(cfg["LR"], 1e-20, total, i)
        for u in range(nupd):
            k = jr.split(rng, 2); carry = k[1]                             # :222
            new = []
            for _ in range(T):
                (hs, obs, ld, la, st, carry), tr, info = _oracle_step(env, params, hs, obs, ld, la, st, carry, 1.0, 0.1, E)
                new.append(tr)
            rng = carry
            mem = mem[T:] + new                                            # :239-243
            stack = {kk: np.stack([m[kk] for m in mem]) for kk in mem[0]}  # [Tm, E, ...]
            k = jr.split(rng, 2); r = k[0]                                 # :381
            losses = []
            for _ in range(cfg["NUM_EPOCHS"]):
                k = jr.split(r, 2); r, kperm = k[0], k[1]                  # :368
                perm = jr.permutation_indices(kperm, E)
                r = jr.split(r, 2)[0]                                      # :375
                for mb in range(nmb):
                    idx = perm[mb * Bm:(mb + 1) * Bm]
                    loss, chosen, g = RR.rnn_loss_and_grads(
                        params, stack["last_hs"][0][idx], stack["obs"][:, idx], stack["last_done"][:, idx],
                        stack["last_action"][:, idx], stack["action"][:, idx], stack["reward"][:, idx],
                        stack["done"][:, idx], cfg["GAMMA"], cfg["LAMBDA"])
                    params, opt, _ = R.radam_clip_step(params, g, opt, lr_fn(opt["count"]), cfg["MAX_GRAD_NORM"])
                    losses.append(loss)
            rng = r
            got = float(out["metrics"]["td_loss"][s, u])
            assert abs(got - np.mean(losses)) < 2e-3 * max(1.0, abs(np.mean(losses))), (u, got, np.mean(losses))
        for p, *_ in eng.spec.entries:
            d = np.abs(leaf(ts.params, p) - params["/".join(p)])
            assert np.quantile(d, 0.99) < 1e-4 and d.max() < 1e-3, (p, d.max())
        assert np.array_equal(out["runner_state"][4][s].cpu().numpy().view(np.uint32), rng)


def test_rnn_cartpole_smoke_with_eval():
    from purejaxql_b200 import config_loader, pqn_rnn_gymnax
    c = config_loader.compose(["+alg=pqn_rnn_cartpole", "NUM_SEEDS=2", "SAVE_PATH=null", "alg.TOTAL_TIMESTEPS=16384",
                               "alg.TOTAL_TIMESTEPS_DECAY=16384", "alg.TEST_NUM_ENVS=16", "alg.TEST_INTERVAL=0.5",
                               "alg.HIDDEN_SIZE=128"])
    cfg = {**c, **c["alg"]}
    out = pqn_rnn_gymnax.make_train(cfg)(jr.split(jr.PRNGKey(0), 2))
    m = out["metrics"]
    assert m["td_loss"].shape == (2, 8) and torch.isfinite(m["td_loss"]).all()
    assert "test/returned_episode_returns" in m and "env_frame" not in m
    assert m["env_step"][0, -1].item() == 16384


def test_rnn_cuda_graph_replay_equals_eager():
    """The recurrent update is captured into a CUDA graph after the first eager update (these runs are launch-bound);
    replayed updates must reproduce the eager run bit for bit (everything is deterministic)."""
    from purejaxql_b200 import pqn_rnn_gymnax
    outs = []
    for graph in (False, True):
        cfg = dict(ENV_NAME="CartPole-v1", NUM_ENVS=8, NUM_STEPS=12, MEMORY_WINDOW=3, NUM_MINIBATCHES=4, NUM_EPOCHS=2,
                   EPS_START=1.0, EPS_FINISH=0.1, EPS_DECAY=0.5, LR=1e-4, MAX_GRAD_NORM=10, GAMMA=0.99, LAMBDA=0.95,
                   NORM_TYPE="layer_norm", NORM_INPUT=False, HIDDEN_SIZE=128, NUM_LAYERS=2, LR_LINEAR_DECAY=True,
                   REW_SCALE=0.1, WANDB_MODE="disabled", TEST_DURING_TRAINING=True, TEST_INTERVAL=0.4, TEST_NUM_ENVS=8,
                   EPS_TEST=0.0, CUDA_GRAPH=graph)
        cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(5 * cfg["NUM_STEPS"] * cfg["NUM_ENVS"])
        train = pqn_rnn_gymnax.make_train(cfg)
        out = train(jr.split(jr.PRNGKey(5), 2))
        assert train.engine.graph_captured == graph
        outs.append((out["runner_state"][0].params_flat.cpu().numpy(), out["metrics"]["td_loss"].cpu().numpy(),
                     out["metrics"]["test/returned_episode_returns"].cpu().numpy(),
                     out["runner_state"][4].cpu().numpy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b, equal_nan=True)
