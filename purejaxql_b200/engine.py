"""The PQN training program: ``make_train(config) -> train(rngs)``.

Host-side restatement of ``make_train`` in purejaxql/pqn_minatar.py:89-431 and
purejaxql/pqn_gymnax.py:78-424 (identical modulo the network and the flatten
wrapper), with the seed axis taken natively — the reference wraps ``train`` in
``jax.jit(jax.vmap(...))`` over ``rngs[S,2]`` (pqn_minatar.py:459-461); here
``train(rngs)`` receives the whole ``[S,2]`` key array and every kernel launch
covers all S seeds.

All compute is libpqn_b200 kernels (include/pqn_b200.h).  This module only
allocates buffers (torch), walks the reference's PRNG key chain (SURVEY
Appendix B) and sequences the launches:

  per update (``_update_step``, :176-369):
    rollout   T x [ Q-network forward ; fused eps-greedy + env step + stores ]
    bootstrap forward on the last obs ; Q(lambda) reverse scan
    epochs x minibatches x [ permutation gather ; loss/grad ; clip+RAdam ; BN stats ]
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib, envs, jaxrandom as jr
from .networks import NET_CNN, NET_MLP, QNetworkSpec

INFO_KEYS = ("returned_episode_returns", "returned_episode_lengths", "timestep", "returned_episode", "discount")


def _f32(x):
    return np.float32(x)


def linear_schedule(init, end, transition_steps, count):
    """optax.linear_schedule evaluated in float32 (pqn_minatar.py:134-146)."""
    if transition_steps <= 0:
        return _f32(init)
    c = np.clip(_f32(count), _f32(0), _f32(transition_steps))
    frac = _f32(1) - c / _f32(transition_steps)
    return _f32(_f32(init - end) * frac + _f32(end))


def radam_schedule_table(num_steps, lr_fn, b1=0.9, b2=0.999, threshold=5.0):
    """[num_steps,4] float32 rows (lr_t, 1-b1^t, 1-b2^t, rect_t or 0) of
    optax.scale_by_radam + scale_by_learning_rate for optimizer steps t=1..

    Evaluated in float32 in optax's operation order (the traced program computes ``b2t = b2**count_inc``,
    ``ro = ro_inf - 2*count_inc*b2t/(1-b2t)`` and the rectification term on weak-typed float32 scalars): the
    cancellation in ``ro`` is visible in float32 for the first few hundred steps, so a float64 table would not
    match the reference optimizer to 1e-5 there.  (XLA's f32 pow may still differ from numpy's by an ulp, which
    the cancellation amplifies; this cannot be pinned without an optax install -- DESIGN.md section 5.)"""
    f = np.float32
    tab = np.zeros((max(num_steps, 1), 4), np.float32)
    ro_inf = f(2.0) / (f(1.0) - f(b2)) - f(1.0)
    for i in range(num_steps):
        t = f(i + 1)
        b2t = np.power(f(b2), t, dtype=np.float32)
        b1t = np.power(f(b1), t, dtype=np.float32)
        ro = ro_inf - f(2.0) * t * b2t / (f(1.0) - b2t)
        rect = f(0.0)
        if ro >= f(threshold):
            rect = np.sqrt((ro - f(4)) * (ro - f(2)) * ro_inf / ((ro_inf - f(4)) * (ro_inf - f(2)) * ro), dtype=np.float32)
        tab[i] = (lr_fn(i), f(1.0) - b1t, f(1.0) - b2t, rect)
    return tab


class TrainState(SimpleNamespace):
    """Mirror of CustomTrainState (pqn_minatar.py:82-86): params / batch_stats
    nested dicts with a leading seed axis, plus timesteps, n_updates, grad_steps
    and the optimizer moments."""


class PQNEngine:
    def __init__(self, config: dict, network: str, flatten_obs: bool, device=None):
        self.cfg = config
        self.device = torch.device(device or "cuda")
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.PqnError("purejaxql_b200 needs a CUDA device: there is no CPU fallback")
        _lib.lib()
        c = config
        norm_type, norm_input = c.get("NORM_TYPE", "layer_norm"), bool(c.get("NORM_INPUT", False))
        self.rng_mode = int(c.get("JAX_THREEFRY_PARTITIONABLE", 0))
        self.env, self.env_params = envs.make(c["ENV_NAME"], flatten_obs=flatten_obs, rng_mode=self.rng_mode)
        self.max_steps = int(self.env_params.max_steps_in_episode)
        self.T = int(c["NUM_STEPS"])
        self.E = int(c["NUM_ENVS"])
        self.NU = int(c["NUM_UPDATES"])
        self.A = self.env.num_actions
        self.binary = self.env.binary_obs
        self.network = network
        if network == "cnn":
            if not self.binary:
                raise ValueError("the MinAtar CNN needs a (10,10,C) binary-observation env")
            self.spec = QNetworkSpec(NET_CNN, self.env.info.obs_shape[2], self.A, norm_type=norm_type,
                                     norm_input=norm_input)
            self.row_words = self.env.packed_obs_words          # int32 words per obs row
            self.obs_dtype = torch.int32
        else:
            self.spec = QNetworkSpec(NET_MLP, self.env.obs_dim, self.A, int(c.get("HIDDEN_SIZE", 128)),
                                     int(c.get("NUM_LAYERS", 2)), norm_type=norm_type, norm_input=norm_input)
            if self.binary:
                raise NotImplementedError("MLP on packed MinAtar observations is not built")
            self.row_words = self.env.obs_dim
            self.obs_dtype = torch.float32
        self.nmb = int(c["NUM_MINIBATCHES"])
        self.epochs = int(c["NUM_EPOCHS"])
        self.mb = self.T * self.E // self.nmb
        self.gamma = float(c["GAMMA"])
        self.lam = float(c["LAMBDA"])
        self.rew_scale = float(c.get("REW_SCALE", 1))
        self.test = bool(c.get("TEST_DURING_TRAINING", False))
        self._ws = None

    # ------------------------------------------------------------------ #
    def _workspace(self, S, rows):
        need = int(_lib.lib().pqn_net_workspace_bytes(self.spec.desc, S, rows))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, params, obs, S, rows, obs_rows_per_seed, q_out, gather=None, batch_stats=None):
        """network.apply({"params", "batch_stats"}, obs, train=False) (pqn_minatar.py:184-191)."""
        ws = self._workspace(S, rows)
        if batch_stats is None:
            batch_stats = getattr(self, "_cur_stats", None)
        _lib.check(_lib.lib().pqn_qnet_forward(self.spec.desc, _lib.p(params), _lib.p(batch_stats), _lib.raw(obs),
                                               _lib.p(gather),
                                               obs_rows_per_seed, _lib.p(q_out), S, rows, _lib.p(ws),
                                               _lib.stream_ptr()), "pqn_qnet_forward")
        return q_out

    # ------------------------------------------------------------------ #
    def train(self, rngs):
        c, dev, L = self.cfg, self.device, _lib.lib()
        T, E, A, NU = self.T, self.E, self.A, self.NU
        keys = jr.as_key_tensor(rngs, dev)
        assert keys.dim() == 2 and keys.shape[1] == 2, "train(rngs) takes the [NUM_SEEDS, 2] key array"
        S = keys.shape[0]
        mode = self.rng_mode
        # ---- env-sharded data parallelism (needed when NUM_SEEDS < #GPUs).  Rank r owns envs
        # [r*E/W, (r+1)*E/W) of EVERY seed; rollouts are local (per-env keys are those of the unsharded vmap), each
        # minibatch step all-reduces (mean) the flat [S][P] gradient once before clip + RAdam, so parameters stay
        # bit-identical across ranks.  The minibatch permutation is per rank (statistically equivalent to the
        # reference's global shuffle, not sample-identical).
        shard = getattr(self, "env_shard", None)
        E_total, env_lo, world, rank = E, 0, 1, 0
        if shard is not None and shard[1] > 1:
            import torch.distributed as dist
            rank, world = int(shard[0]), int(shard[1])
            assert E % world == 0, f"NUM_ENVS={E} must be divisible by the {world} env shards"
            E = E // world
            env_lo = rank * E
            assert (T * E) % self.nmb == 0, "NUM_MINIBATCHES must divide NUM_STEPS * NUM_ENVS / world"
            if self.spec.norm_type == "batch_norm" or self.spec.norm_input:
                raise NotImplementedError("env-sharded data parallelism with batch statistics on the path "
                                          "(NORM_TYPE=batch_norm / NORM_INPUT) would need their all-reduce; use "
                                          "DATA_PARALLEL=seeds")
        mb = T * E // self.nmb                                        # minibatch rows of THIS rank
        spec, P = self.spec, self.spec.total
        W = self.row_words

        # ---- schedules (pqn_minatar.py:134-147)
        nud = c["NUM_UPDATES_DECAY"]
        eps_table = torch.tensor([linear_schedule(c["EPS_START"], c["EPS_FINISH"], c["EPS_DECAY"] * nud, n)
                                  for n in range(max(NU, 1))], dtype=torch.float32, device=dev)
        total_grad_steps = NU * self.nmb * self.epochs
        if c.get("LR_LINEAR_DECAY", False):
            lr_fn = lambda i: linear_schedule(c["LR"], 1e-20, nud * self.nmb * self.epochs, i)
        else:
            lr_fn = lambda i: _f32(c["LR"])
        sched = torch.from_numpy(radam_schedule_table(total_grad_steps, lr_fn)).to(dev)

        # ---- key chain (pqn_minatar.py:172-173,415-423)
        k = jr.split(keys, 2, mode)
        K1 = k[:, 0].contiguous()                                   # :172 rng (also the init key, :173)
        params = spec.init(K1, dev)                                 # :156-170
        mu = torch.zeros_like(params)
        nu = torch.zeros_like(params)
        grads = torch.zeros_like(params)
        F = spec.in_c
        batch_stats = spec.init_stats(S, dev)                       # flax BatchNorm running statistics: mean 0, var 1
        self._cur_stats = batch_stats                               # read by forward() (train=False => running stats)
        bn_sums = torch.zeros(S, 2 * F, device=dev)
        step_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        gnorm = torch.zeros(S * 64, device=dev)                     # block partials of the squared gradient norm

        k = jr.split(K1, 2, mode)
        K2, kT0 = k[:, 0].contiguous(), k[:, 1].contiguous()        # :415
        test_metrics = self.get_test_metrics(params, kT0) if self.test else None
        k = jr.split(K2, 2, mode)
        K3, kR = k[:, 0].contiguous(), k[:, 1].contiguous()         # :418
        # ---- rollout buffers: obs rows [S][T+1][E], transitions [S][T][E]
        obs_buf = torch.zeros((S, T + 1, E, W), dtype=self.obs_dtype, device=dev)
        act_buf = torch.zeros((S, T, E), dtype=torch.int32, device=dev)
        rew_buf = torch.zeros((S, T, E), dtype=torch.float32, device=dev)
        done_buf = torch.zeros((S, T, E), dtype=torch.uint8, device=dev)
        maxq_buf = torch.zeros((S, T, E), dtype=torch.float32, device=dev)
        targets = torch.zeros((S, T, E), dtype=torch.float32, device=dev)
        q_buf = torch.zeros((S * E, A), dtype=torch.float32, device=dev)
        info_sums = torch.zeros((S, 5), dtype=torch.float64, device=dev)
        loss_sum = torch.zeros(S, device=dev)
        qsa_sum = torch.zeros(S, device=dev)
        step_keys = torch.zeros((T, S, 2, 2), dtype=torch.int32, device=dev)
        eps_dev = torch.zeros(1, device=dev)
        # ---- reset (vmap_reset, :107-109,419)
        reset_keys = jr.split(kR, E_total, mode)[:, env_lo:env_lo + E].reshape(S * E, 2).contiguous()
        state = torch.empty((self.env.state_words, S * E), dtype=torch.int32, device=dev)
        _lib.check(L.pqn_env_reset(self.env.env_id, _lib.p(reset_keys), _lib.p(state), None, S * E, self.max_steps,
                                   mode, _lib.stream_ptr()), "pqn_env_reset")
        self._write_obs(state, obs_buf, T, S)                        # update_body moves row T to row 0
        rng = jr.split(K3, 2, mode)[:, 1].contiguous()              # :422-423 runner rng

        # pqn_minatar.py:330-338 reports env_frame; pqn_gymnax.py:324-331 does not
        metric_names = ["env_step", "update_steps", *(["env_frame"] if self.network == "cnn" else []), "grad_steps",
                        "td_loss", "qvals", *INFO_KEYS]
        metrics = {m: torch.zeros((S, max(NU, 1)), dtype=torch.float64, device=dev) for m in metric_names}
        test_hist = None
        test_every = None
        if self.test:
            test_every = int(NU * c["TEST_INTERVAL"])
            test_hist = {kk: torch.zeros((S, max(NU, 1)), dtype=torch.float64, device=dev) for kk in INFO_KEYS}
        obs_channels = self.env.observation_space().shape[-1]
        sp = _lib.stream_ptr
        timesteps = 0
        grad_steps = 0
        seed_stride_obs = (T + 1) * E
        seed_stride_tr = T * E
        # ---- static buffers of the update step (the whole step is CUDA-graph capturable)
        rng_buf = rng.clone()                                        # runner rng, updated in place
        kT_buf = torch.zeros((S, 2), dtype=torch.int32, device=dev)  # eval key of this update (:341)
        upd_idx = torch.zeros(1, dtype=torch.int64, device=dev)      # n_updates on the device
        m_cur = torch.zeros((S, 7), dtype=torch.float64, device=dev)  # td_loss, qvals, 5 info means
        ws = self._workspace(S, max(mb, E))
        perm_ws = jr.permutation_workspace(T * E, S, dev)
        denom = float(self.epochs * self.nmb)
        bn_count = float(mb * world * (100 if self.binary else 1))

        def allreduce_(t, avg):
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)              # in-stream: the consumer kernels just follow
                if avg:
                    t.mul_(1.0 / world)

        def update_body():
            """One `_update_step` (pqn_minatar.py:176-350) on the current stream; reads/writes only the static
            buffers above, so it can be replayed from a CUDA graph."""
            # ================= SAMPLE PHASE (:181-219)
            eps_dev.copy_(eps_table.index_select(0, upd_idx))
            obs_buf[:, 0].copy_(obs_buf[:, T])                       # last_obs of this rollout = last next_obs
            carry = jr.split(rng_buf, 2, mode)[:, 1].contiguous()    # :213  `_rng`
            _lib.check(L.pqn_rollout_keys(_lib.p(carry), _lib.p(step_keys), S, T, mode, sp()), "pqn_rollout_keys")
            info_sums.zero_()
            for t in range(T):
                self.forward(params, obs_buf[:, t], S, E, seed_stride_obs, q_buf)
                _lib.check(L.pqn_rollout_act_step(
                    self.env.env_id, _lib.p(step_keys[t]), _lib.p(q_buf), _lib.p(eps_dev), _lib.p(state),
                    _lib.raw(obs_buf[:, t + 1]), seed_stride_obs,
                    _lib.raw(act_buf[:, t]), _lib.raw(rew_buf[:, t]),
                    _lib.raw(done_buf[:, t]), _lib.raw(maxq_buf[:, t]),
                    seed_stride_tr, _lib.p(info_sums), 0, S, E, E_total, env_lo, self.max_steps, self.rew_scale, mode,
                    sp()), "pqn_rollout_act_step")
            r = carry                                                # scan's final carry (:214)
            # ================= bootstrap + Q(lambda) (:227-260)
            self.forward(params, obs_buf[:, T], S, E, seed_stride_obs, q_buf)
            _lib.check(L.pqn_qlambda(_lib.p(rew_buf), _lib.p(done_buf), _lib.p(maxq_buf), _lib.p(q_buf),
                                     _lib.p(targets), T, S, E, A, self.gamma, self.lam, sp()), "pqn_qlambda")
            # ================= NETWORKS UPDATE (:263-327)
            r = jr.split(r, 2, mode)[:, 0].contiguous()              # :324
            loss_sum.zero_()
            qsa_sum.zero_()
            for _ in range(self.epochs):
                k = jr.split(r, 2, mode)                             # :309
                r, kperm = k[:, 0].contiguous(), k[:, 1].contiguous()
                if world > 1:                                        # a different local permutation on every rank
                    kperm = jr.split(kperm, world, mode)[:, rank].contiguous()
                perm_view = jr.permutation_indices(kperm, T * E, mode, chunk=mb, workspace=perm_ws)   # :299-321  [nmb][S][mb]
                r = jr.split(r, 2, mode)[:, 0].contiguous()          # :317
                for mbi in range(self.nmb):
                    _lib.check(L.pqn_qnet_loss_grad(
                        spec.desc, _lib.p(params), _lib.p(batch_stats), _lib.p(obs_buf), _lib.p(perm_view[mbi]),
                        seed_stride_obs,
                        _lib.p(act_buf), _lib.p(targets), seed_stride_tr, _lib.p(grads), _lib.p(loss_sum),
                        _lib.p(qsa_sum), _lib.p(bn_sums), S, mb, _lib.p(ws), sp()), "pqn_qnet_loss_grad")
                    allreduce_(grads, True)                          # the ONE collective of the data path
                    allreduce_(bn_sums, False)
                    _lib.check(L.pqn_radam_clip_step(_lib.p(params), _lib.p(grads), _lib.p(mu), _lib.p(nu),
                                                     _lib.p(sched), _lib.p(step_counter), _lib.p(gnorm), S, P,
                                                     float(c["MAX_GRAD_NORM"]), 0.9, 0.999, 1e-8, sp()),
                               "pqn_radam_clip_step")
                    _lib.check(L.pqn_bn_stats_update(_lib.p(batch_stats), _lib.p(bn_sums), S, F, spec.stats_total,
                                                     bn_count, 0.99, sp()), "pqn_bn_stats_update")
            if self.test:                                            # :341  rng, _rng = split(rng)
                k = jr.split(r, 2, mode)
                r = k[:, 0].contiguous()
                kT_buf.copy_(k[:, 1])
            rng_buf.copy_(r)
            allreduce_(loss_sum, True)
            allreduce_(qsa_sum, True)
            allreduce_(info_sums, False)
            m_cur[:, 0] = loss_sum.double() / denom
            m_cur[:, 1] = qsa_sum.double() / denom
            m_cur[:, 2:7] = info_sums / float(T * E_total)
            upd_idx.add_(1)

        # CUDA graph: "auto" captures the update when the run is launch-bound (small S*E); the first update runs
        # eagerly (warms every code path), later updates replay the captured graph.
        want_graph = c.get("CUDA_GRAPH", "auto")
        use_graph = (S * E * T <= (1 << 21) and world == 1) if want_graph == "auto" else bool(want_graph)
        use_graph = use_graph and NU > 2
        graph = None
        self.graph_captured = False
        self.graph_replays = 0
        self.graph_launches_per_replay = 0

        # hooks (bench / tests): both are called on the host OUTSIDE the captured region, so they do not prevent
        # CUDA-graph replay.  on_update_end receives the static rollout buffers of the update that just ran.
        on_update_begin = getattr(self, "on_update_begin", None)
        on_update_end = getattr(self, "on_update_end", None)
        dbg = dict(obs=obs_buf, action=act_buf, reward=rew_buf, done=done_buf, maxq=maxq_buf, targets=targets,
                   params=params, state=state, rng=rng_buf)
        for n_updates in range(NU):
            if on_update_begin is not None:
                on_update_begin(n_updates)
            if graph is not None:
                graph.replay()
                self.graph_replays += 1
            else:
                update_body()
                if use_graph and n_updates == 0:
                    try:
                        torch.cuda.synchronize(dev)
                        g = torch.cuda.CUDAGraph()
                        l0 = L.pqn_launch_count()
                        with torch.cuda.graph(g):
                            update_body()
                        self.graph_launches_per_replay = int(L.pqn_launch_count() - l0)
                        graph = g
                        self.graph_captured = True
                    except Exception as e:                            # capture is an optimisation only
                        import warnings
                        warnings.warn(f"CUDA graph capture of the update step failed ({e!r}); running eagerly")
                        graph = None
                        use_graph = False
                        torch.cuda.synchronize(dev)
            if on_update_end is not None:
                on_update_end(n_updates, dbg)
            timesteps += T * E_total                                 # :222-225
            grad_steps += self.nmb * self.epochs
            # ================= metrics (:329-338)
            n_done = n_updates + 1
            col = n_updates
            metrics["env_step"][:, col] = timesteps
            metrics["update_steps"][:, col] = n_done
            if "env_frame" in metrics:
                metrics["env_frame"][:, col] = timesteps * obs_channels
            metrics["grad_steps"][:, col] = grad_steps
            metrics["td_loss"][:, col] = m_cur[:, 0]
            metrics["qvals"][:, col] = m_cur[:, 1]
            for j, kk in enumerate(INFO_KEYS):
                metrics[kk][:, col] = m_cur[:, 2 + j]
            # ================= evaluation (:340-350)
            if self.test:
                if test_every > 0 and n_done % test_every == 0:
                    test_metrics = self.get_test_metrics(params, kT_buf.clone())
                for kk in INFO_KEYS:
                    test_hist[kk][:, col] = test_metrics[kk]
            if c.get("WANDB_MODE", "disabled") != "disabled":
                self._wandb_log(metrics, test_hist, col, jr.to_numpy_u32(keys)[:, 0])
        rng = rng_buf

        torch.cuda.synchronize(dev)
        out_metrics = {m: v[:, :NU].float() if m in ("td_loss", "qvals", *INFO_KEYS) else v[:, :NU].to(torch.int64)
                       for m, v in metrics.items()}
        if self.test:
            out_metrics.update({f"test/{kk}": v[:, :NU].float() for kk, v in test_hist.items()})
        train_state = TrainState(
            params=spec.unflatten(params), params_flat=params,
            batch_stats=spec.unflatten_stats(batch_stats), batch_stats_flat=batch_stats,
            opt_state=SimpleNamespace(mu=mu, nu=nu, count=grad_steps),
            timesteps=torch.full((S,), timesteps, dtype=torch.int64), n_updates=torch.full((S,), NU),
            grad_steps=torch.full((S,), grad_steps))
        expl_state = (self._final_obs(obs_buf, S), state)
        return {"runner_state": (train_state, expl_state, test_metrics, rng), "metrics": out_metrics}

    # ------------------------------------------------------------------ #
    def _write_obs(self, state, obs_buf, t, S):
        """obs rows of `state` into obs_buf[:, t] (the reset observation)."""
        L = _lib.lib()
        E = obs_buf.shape[2]
        tmp = torch.empty((S * E, self.row_words), dtype=self.obs_dtype, device=self.device)
        if self.binary:
            _lib.check(L.pqn_env_obs_packed(self.env.env_id, _lib.p(state), _lib.p(tmp), S * E, _lib.stream_ptr()),
                       "pqn_env_obs_packed")
        else:
            _lib.check(L.pqn_env_obs(self.env.env_id, _lib.p(state), _lib.p(tmp), S * E, _lib.stream_ptr()),
                       "pqn_env_obs")
        obs_buf[:, t] = tmp.view(S, E, self.row_words)

    def _final_obs(self, obs_buf, S):
        st_obs = obs_buf[:, -1]
        return st_obs.contiguous()

    # ------------------------------------------------------------------ #
    def get_test_metrics(self, params, rng):
        """Greedy evaluation rollout (pqn_minatar.py:371-413), incl. its key quirks:
        the scan carry starts at the reset key `_rng`, and each step uses the same
        sub-key for the action keys and the env keys."""
        c, dev, L, mode = self.cfg, self.device, _lib.lib(), self.rng_mode
        S = rng.shape[0]
        N = int(c["TEST_NUM_ENVS"])
        steps = int(c["TEST_NUM_STEPS"])
        W, A = self.row_words, self.A
        k = jr.split(rng, 2, mode)
        kr = k[:, 1].contiguous()                                    # :396 `_rng`
        state = torch.empty((self.env.state_words, S * N), dtype=torch.int32, device=dev)
        _lib.check(L.pqn_env_reset(self.env.env_id, _lib.p(jr.split(kr, N, mode).reshape(S * N, 2).contiguous()),
                                   _lib.p(state), None, S * N, self.max_steps, mode, _lib.stream_ptr()),
                   "pqn_env_reset")
        obs = torch.zeros((S, 2, N, W), dtype=self.obs_dtype, device=dev)   # ping-pong rows
        self._write_obs(state, obs, 0, S)
        q = torch.zeros((S * N, A), dtype=torch.float32, device=dev)
        scratch_i = torch.zeros((S, N), dtype=torch.int32, device=dev)
        scratch_f = torch.zeros((S, N), dtype=torch.float32, device=dev)
        scratch_f2 = torch.zeros((S, N), dtype=torch.float32, device=dev)
        scratch_b = torch.zeros((S, N), dtype=torch.uint8, device=dev)
        sums = torch.zeros((S, 5), dtype=torch.float64, device=dev)
        eps = torch.full((1,), float(c["EPS_TEST"]), device=dev)
        carry = kr                                                   # :399-401
        for t in range(steps):
            k = jr.split(carry, 2, mode)                             # :378
            carry, ku = k[:, 0].contiguous(), k[:, 1].contiguous()
            sk = torch.stack([ku, ku], 1).contiguous()               # same key for actions and env (:388-393)
            cur, nxt = t & 1, (t + 1) & 1
            self.forward(params, obs[:, cur], S, N, 2 * N, q)
            _lib.check(L.pqn_rollout_act_step(
                self.env.env_id, _lib.p(sk), _lib.p(q), _lib.p(eps), _lib.p(state),
                _lib.raw(obs[:, nxt]), 2 * N, _lib.p(scratch_i), _lib.p(scratch_f), _lib.p(scratch_b),
                _lib.p(scratch_f2), N, _lib.p(sums), 1, S, N, 0, 0, self.max_steps, 1.0, mode, _lib.stream_ptr()),
                "pqn_rollout_act_step")
        cnt = sums[:, 3]
        out = {}
        for j, kk in enumerate(INFO_KEYS):
            out[kk] = torch.where(cnt > 0, sums[:, j] / cnt.clamp(min=1), torch.full_like(cnt, float("nan")))
        return out

    def _wandb_log(self, metrics, test_hist, col, seed_labels):
        import wandb
        S = metrics["td_loss"].shape[0]
        row = {m: v[:, col].mean().item() for m, v in metrics.items()}
        if test_hist is not None:
            row.update({f"test/{kk}": v[:, col].nanmean().item() for kk, v in test_hist.items()})
        if self.cfg.get("WANDB_LOG_ALL_SEEDS", False):
            for s in range(S):
                for m, v in metrics.items():
                    row[f"rng{int(seed_labels[s])}/{m}"] = v[s, col].item()
        wandb.log(row, step=int(row["update_steps"]))


def prepare_config(config: dict, env_max_steps: int, allow_test_steps_override: bool):
    """The config mutations of make_train (pqn_minatar.py:91-105 / pqn_gymnax.py:80-97)."""
    config["NUM_UPDATES"] = config["TOTAL_TIMESTEPS"] // config["NUM_STEPS"] // config["NUM_ENVS"]
    config["NUM_UPDATES_DECAY"] = config["TOTAL_TIMESTEPS_DECAY"] // config["NUM_STEPS"] // config["NUM_ENVS"]
    assert (config["NUM_STEPS"] * config["NUM_ENVS"]) % config["NUM_MINIBATCHES"] == 0, \
        "NUM_MINIBATCHES must divide NUM_STEPS*NUM_ENVS"
    if allow_test_steps_override:
        config["TEST_NUM_STEPS"] = config.get("TEST_NUM_STEPS", env_max_steps)
    else:
        config["TEST_NUM_STEPS"] = env_max_steps
    return config
