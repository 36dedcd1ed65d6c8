"""GPU end-to-end parity: one full `_update_step` of the engine against the
oracle's `update_step` (same weights, same keys => same rollout, same
permutation, same minibatches), and learning smoke tests."""
import numpy as np
import pytest
import torch

from oracle import gymnax_envs as G
from oracle import jax_prng as jr
from oracle import pqn_ref as R

pytestmark = pytest.mark.gpu


def _cfg(env, **kw):
    c = dict(ENV_NAME=env, TOTAL_TIMESTEPS=0, TOTAL_TIMESTEPS_DECAY=0, NUM_ENVS=64, NUM_STEPS=8, NUM_MINIBATCHES=4,
             NUM_EPOCHS=2, EPS_START=1.0, EPS_FINISH=0.05, EPS_DECAY=0.1, LR=5e-4, MAX_GRAD_NORM=10, GAMMA=0.99,
             LAMBDA=0.65, NORM_TYPE="layer_norm", LR_LINEAR_DECAY=True, WANDB_MODE="disabled",
             TEST_DURING_TRAINING=False)
    c.update(kw)
    return c


def _run_updates_against_oracle(module, env_name, kind, flatten, cfg, S=2, nupd=3, graph=False):
    """eps = 1 for the whole run (every action random => every rollout is integer-exact), so ALL updates of the
    engine — obs hand-over between updates, key chain, permutations, optimizer — must track the oracle."""
    cfg["EPS_START"] = cfg["EPS_FINISH"] = 1.0
    cfg["CUDA_GRAPH"] = graph
    cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(nupd * cfg["NUM_STEPS"] * cfg["NUM_ENVS"])
    train = module.make_train(cfg)
    eng = train.engine
    rngs = jr.split(jr.PRNGKey(0), S)
    captured = {}
    orig_init = eng.spec.init

    def init(keys, device):
        flat = orig_init(keys, device)
        captured["flat"] = flat.clone()
        return flat
    eng.spec.init = init
    out = train(rngs)
    assert eng.graph_captured == bool(graph)
    ts = out["runner_state"][0]
    tree0 = eng.spec.unflatten(captured["flat"])
    T, E = cfg["NUM_STEPS"], cfg["NUM_ENVS"]
    for s in range(S):
        def leaf(tree, path, s=s):
            d = tree
            for k in path:
                d = d[k]
            return d[s].cpu().numpy()
        params = {"/".join(p): leaf(tree0, p).astype(np.float32) for p, *_ in eng.spec.entries}
        K1 = jr.split(rngs[s], 2)[0]                                 # oracle key chain
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
        for u in range(nupd):
            params, opt, bs, obs, st, rng, m, tr, tg = R.update_step(env, kind, params, opt, bs, obs, st, rng, dict(cfg),
                                                                     u, lr_fn)
            got = {kk: float(v[s, u]) for kk, v in out["metrics"].items()}
            # integer games: exact rollouts; fp32 physics (classic control) may flip a termination once in a while
            itol, ltol = (1e-6, 1e-4) if kind == "cnn" else (2e-2, 2e-2)
            for kk in ("returned_episode_returns", "returned_episode_lengths", "timestep", "returned_episode",
                       "discount"):
                assert abs(got[kk] - m[kk]) < itol * max(1, abs(m[kk])), (u, kk, got[kk], m[kk])
            assert abs(got["td_loss"] - m["td_loss"]) < ltol * max(1.0, abs(m["td_loss"])), (u, got["td_loss"], m["td_loss"])
            assert abs(got["qvals"] - m["qvals"]) < ltol * max(1.0, abs(m["qvals"])), u
        if kind == "cnn":
            for p, *_ in eng.spec.entries:                           # parameters after ALL updates
                assert np.abs(leaf(ts.params, p) - params["/".join(p)]).max() < 5e-5, p
        assert np.array_equal(out["runner_state"][3][s].cpu().numpy().view(np.uint32), rng)
    return out


def test_minatar_update_step_matches_oracle():
    from purejaxql_b200 import pqn_minatar
    cfg = _cfg("Breakout-MinAtar")
    _run_updates_against_oracle(pqn_minatar, "Breakout-MinAtar", "cnn", False, cfg)


def test_minatar_update_steps_under_cuda_graph_match_oracle():
    from purejaxql_b200 import pqn_minatar
    cfg = _cfg("Breakout-MinAtar")
    _run_updates_against_oracle(pqn_minatar, "Breakout-MinAtar", "cnn", False, cfg, nupd=4, graph=True)


def test_gymnax_update_step_matches_oracle():
    from purejaxql_b200 import pqn_gymnax
    cfg = _cfg("CartPole-v1", HIDDEN_SIZE=128, NUM_LAYERS=2, REW_SCALE=0.1, LAMBDA=0.95, NUM_ENVS=32, NUM_STEPS=16)
    _run_updates_against_oracle(pqn_gymnax, "CartPole-v1", "mlp", True, cfg)


def test_gymnax_update_steps_under_cuda_graph_match_oracle():
    from purejaxql_b200 import pqn_gymnax
    cfg = _cfg("CartPole-v1", HIDDEN_SIZE=128, NUM_LAYERS=2, REW_SCALE=0.1, LAMBDA=0.95, NUM_ENVS=32, NUM_STEPS=16)
    _run_updates_against_oracle(pqn_gymnax, "CartPole-v1", "mlp", True, cfg, nupd=4, graph=True)


def test_params_after_first_update_match_oracle():
    """Single update, eps=1 (all actions random => rollout bit-exact): parameters after
    the NUM_EPOCHS x NUM_MINIBATCHES clipped-RAdam steps agree with the oracle to 1e-5."""
    from purejaxql_b200 import pqn_minatar
    cfg = _cfg("Breakout-MinAtar", NUM_ENVS=128, NUM_STEPS=8, NUM_MINIBATCHES=4)
    cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(cfg["NUM_STEPS"] * cfg["NUM_ENVS"])
    train = pqn_minatar.make_train(cfg)
    eng = train.engine
    S = 2
    rngs = jr.split(jr.PRNGKey(1), S)
    cap = {}
    orig = eng.spec.init
    eng.spec.init = lambda k, d: cap.setdefault("flat", orig(k, d)).clone()
    out = train(rngs)
    ts = out["runner_state"][0]
    tree0 = eng.spec.unflatten(cap["flat"])
    T, E = cfg["NUM_STEPS"], cfg["NUM_ENVS"]
    for s in range(S):
        def leaf(tree, path):
            d = tree
            for k in path:
                d = d[k]
            return d[s].cpu().numpy()
        params = {"/".join(p): leaf(tree0, p).astype(np.float32) for p, *_ in eng.spec.entries}
        K1 = jr.split(rngs[s], 2)[0]
        K2 = jr.split(K1, 2)[0]
        k = jr.split(K2, 2); K3, kR = k[0], k[1]
        env = G.make("Breakout-MinAtar")
        obs, st = env.reset(jr.split(kR, E))
        rng = jr.split(K3, 2)[1]
        total = cfg["NUM_UPDATES_DECAY"] * cfg["NUM_MINIBATCHES"] * cfg["NUM_EPOCHS"]
        lr_fn = lambda i: R.linear_schedule(cfg["LR"], 1e-20, total, i)
        bs = {"mean": np.zeros(4, np.float32), "var": np.ones(4, np.float32)}
        p2, opt, bs, obs, st, rng2, m, tr, tg = R.update_step(env, "cnn", params, R.opt_init(params), bs, obs, st, rng,
                                                              dict(cfg), 0, lr_fn)
        for p, *_ in eng.spec.entries:
            got = leaf(ts.params, p)
            ref = p2["/".join(p)]
            assert np.abs(got - ref).max() < 1e-5, (p, np.abs(got - ref).max())
        assert np.allclose(ts.batch_stats["BatchNorm_0"]["mean"][s].cpu().numpy(), bs["mean"], atol=1e-6)
        assert np.allclose(ts.batch_stats["BatchNorm_0"]["var"][s].cpu().numpy(), bs["var"], atol=1e-6)
        # final runner rng and env state are integer work: exact
        assert np.array_equal(out["runner_state"][3][s].cpu().numpy().view(np.uint32), rng2)
        from purejaxql_b200 import envs
        f = envs.state_to_fields("Breakout-MinAtar", out["runner_state"][1][1][:, s * E:(s + 1) * E])
        for kk, v in st.items():
            assert np.array_equal(f[kk].cpu().numpy().astype(v.dtype), v), kk


def test_cartpole_learns():
    """PQN on CartPole-v1 with the shipped preset (shortened): mean returned episode
    return of the behaviour policy climbs well above the random-policy ~22."""
    from purejaxql_b200 import config_loader, pqn_gymnax
    c = config_loader.compose(["+alg=pqn_cartpole", "NUM_SEEDS=4", "SAVE_PATH=null", "alg.TOTAL_TIMESTEPS=2e5",
                               "alg.TOTAL_TIMESTEPS_DECAY=2e5", "alg.TEST_DURING_TRAINING=False"])
    cfg = {**c, **c["alg"]}
    train = pqn_gymnax.make_train(cfg)
    out = train(jr.split(jr.PRNGKey(0), 4))
    ret = out["metrics"]["returned_episode_returns"].cpu().numpy()        # [S, NUM_UPDATES]
    assert np.isfinite(out["metrics"]["td_loss"].cpu().numpy()).all()
    assert ret[:, -5:].mean() > 60.0, ret[:, -5:].mean()


def test_minatar_smoke_with_eval_and_save(tmp_path):
    from purejaxql_b200 import config_loader, pqn_minatar
    from purejaxql_b200.utils.save_load import load_params
    c = config_loader.compose(["+alg=pqn_minatar", "alg.ENV_NAME=Breakout-MinAtar", "NUM_SEEDS=2",
                               f"SAVE_PATH={tmp_path}", "alg.TOTAL_TIMESTEPS=4e4", "alg.TOTAL_TIMESTEPS_DECAY=4e4",
                               "alg.NUM_ENVS=64", "alg.TEST_NUM_ENVS=16", "alg.TEST_INTERVAL=0.5"])
    out = pqn_minatar.single_run(c)
    m = out["metrics"]
    assert m["td_loss"].shape == (2, 19) and "test/returned_episode_returns" in m
    assert torch.isfinite(m["td_loss"]).all()
    files = sorted(p.name for p in (tmp_path / "Breakout-MinAtar").iterdir())
    assert "pqn_Breakout-MinAtar_seed0_vmap0.safetensors" in files and "pqn_Breakout-MinAtar_seed0_config.yaml" in files
    tree = load_params(str(tmp_path / "Breakout-MinAtar" / "pqn_Breakout-MinAtar_seed0_vmap1.safetensors"))
    assert tuple(tree["CNN_0"]["Conv_0"]["kernel"].shape) == (3, 3, 4, 16)
    assert tuple(tree["CNN_0"]["Dense_0"]["kernel"].shape) == (1024, 128)
    assert tuple(tree["Dense_0"]["kernel"].shape) == (128, 3) and "BatchNorm_0" in tree


@pytest.mark.parametrize("env_name", ["Asterix-MinAtar", "Freeway-MinAtar", "SpaceInvaders-MinAtar"])
def test_other_minatar_first_update_matches_oracle(env_name):
    """C=7 / C=6 games through the whole engine: one update with eps=1 reproduces the oracle's parameters."""
    from purejaxql_b200 import pqn_minatar
    cfg = _cfg(env_name, NUM_ENVS=64, NUM_STEPS=8, NUM_MINIBATCHES=4)
    cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(cfg["NUM_STEPS"] * cfg["NUM_ENVS"])
    train = pqn_minatar.make_train(cfg)
    eng = train.engine
    rngs = jr.split(jr.PRNGKey(2), 1)
    cap = {}
    orig = eng.spec.init
    eng.spec.init = lambda k, d: cap.setdefault("flat", orig(k, d)).clone()
    out = train(rngs)
    ts = out["runner_state"][0]
    tree0 = eng.spec.unflatten(cap["flat"])
    E = cfg["NUM_ENVS"]

    def leaf(tree, path):
        d = tree
        for k in path:
            d = d[k]
        return d[0].cpu().numpy()
    params = {"/".join(p): leaf(tree0, p).astype(np.float32) for p, *_ in eng.spec.entries}
    K1 = jr.split(rngs[0], 2)[0]
    K2 = jr.split(K1, 2)[0]
    k = jr.split(K2, 2); K3, kR = k[0], k[1]
    env = G.make(env_name)
    obs, st = env.reset(jr.split(kR, E))
    rng = jr.split(K3, 2)[1]
    total = cfg["NUM_UPDATES_DECAY"] * cfg["NUM_MINIBATCHES"] * cfg["NUM_EPOCHS"]
    lr_fn = lambda i: R.linear_schedule(cfg["LR"], 1e-20, total, i)
    C = eng.spec.in_c
    bs = {"mean": np.zeros(C, np.float32), "var": np.ones(C, np.float32)}
    # the oracle's CNN helpers are written for any channel count
    p2, *_ = R.update_step(env, "cnn", params, R.opt_init(params), bs, obs, st, rng, dict(cfg), 0, lr_fn)
    for p, *_ in eng.spec.entries:
        assert np.abs(leaf(ts.params, p) - p2["/".join(p)]).max() < 1e-5, p


def test_wandb_offline_logging_path(tmp_path, monkeypatch):
    """WANDB_MODE != disabled takes the per-update logging branch (pqn_minatar.py:353-365), incl. per-seed keys."""
    pytest.importorskip("wandb")
    import wandb
    from purejaxql_b200 import config_loader, pqn_gymnax
    monkeypatch.setenv("WANDB_DIR", str(tmp_path))
    monkeypatch.setenv("WANDB_SILENT", "true")
    c = config_loader.compose(["+alg=pqn_cartpole", "NUM_SEEDS=2", "SAVE_PATH=null", "WANDB_MODE=offline",
                               "alg.TOTAL_TIMESTEPS=8192", "alg.TOTAL_TIMESTEPS_DECAY=8192", "alg.NUM_ENVS=16",
                               "alg.TEST_NUM_ENVS=8", "alg.TEST_NUM_STEPS=20", "alg.WANDB_LOG_ALL_SEEDS=True",
                               "PROJECT=pqn_b200_test"])
    try:
        out = pqn_gymnax.single_run(c)
    finally:
        wandb.finish()
    assert out["metrics"]["td_loss"].shape == (2, 8)
    assert any(p.name.startswith("offline-run") for p in (tmp_path / "wandb").iterdir())


def test_two_identical_runs_give_bit_identical_parameters():
    """No float atomics on the default path: every cross-row reduction is a fixed-order
    two-stage sum, so train() is run-to-run deterministic like the reference."""
    from purejaxql_b200 import pqn_minatar
    outs = []
    for _ in range(2):
        cfg = _cfg("Breakout-MinAtar", NUM_ENVS=256, NUM_STEPS=8, NUM_MINIBATCHES=4, EPS_START=0.5, EPS_FINISH=0.1,
                   EPS_DECAY=1.0)
        cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(4 * cfg["NUM_STEPS"] * cfg["NUM_ENVS"])
        out = pqn_minatar.make_train(cfg)(jr.split(jr.PRNGKey(4), 3))
        outs.append((out["runner_state"][0].params_flat.cpu().numpy(), out["metrics"]["td_loss"].cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("variant", ["mlp", "mlp_1seed_big", "cnn_batch_norm", "mlp_norm_input", "gru"])
def test_other_network_variants_are_run_to_run_deterministic(variant):
    """The MLP (thin first-layer weight gradient, split-K tensor-core hidden layer at one seed), the modular
    NORM_TYPE / NORM_INPUT path and the GRU engine: every weight gradient is a fixed-order sum as well."""
    from purejaxql_b200 import config_loader, pqn_gymnax, pqn_minatar, pqn_rnn_gymnax
    outs = []
    for _ in range(2):
        if variant == "gru":
            c = config_loader.compose(["+alg=pqn_rnn_cartpole", "NUM_SEEDS=2", "SAVE_PATH=null", "alg.TOTAL_TIMESTEPS=8192",
                                       "alg.TOTAL_TIMESTEPS_DECAY=8192", "alg.TEST_DURING_TRAINING=False",
                                       "alg.HIDDEN_SIZE=128"])
            cfg = {**c, **c["alg"]}
            out = pqn_rnn_gymnax.make_train(cfg)(jr.split(jr.PRNGKey(1), 2))
        elif variant == "cnn_batch_norm":
            cfg = _cfg("Breakout-MinAtar", NUM_ENVS=128, NUM_STEPS=8, NORM_TYPE="batch_norm", NORM_INPUT=False,
                       EPS_START=0.5, EPS_FINISH=0.1, EPS_DECAY=1.0)
            cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(3 * 8 * 128)
            out = pqn_minatar.make_train(cfg)(jr.split(jr.PRNGKey(2), 2))
        else:
            big = variant == "mlp_1seed_big"
            cfg = _cfg("CartPole-v1", NUM_ENVS=4096 if big else 64, NUM_STEPS=16, HIDDEN_SIZE=128, NUM_LAYERS=2,
                       NORM_INPUT=variant == "mlp_norm_input", EPS_START=0.5, EPS_FINISH=0.1, EPS_DECAY=1.0)
            cfg["TOTAL_TIMESTEPS"] = cfg["TOTAL_TIMESTEPS_DECAY"] = float(3 * 16 * cfg["NUM_ENVS"])
            out = pqn_gymnax.make_train(cfg)(jr.split(jr.PRNGKey(3), 1 if big else 3))
        outs.append((out["runner_state"][0].params_flat.cpu().numpy(), out["metrics"]["td_loss"].cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
