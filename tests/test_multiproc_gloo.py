"""world_size-2 gloo test (CPU) of the N>1 path: seeds are independent runs, so
ranks take disjoint contiguous slices of the SAME split(PRNGKey(SEED), NUM_SEEDS);
bench.py's reduction of per-rank times is a MAX all-reduce."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import jax_prng as jr
    from purejaxql_b200 import _runner
    rngs = torch.from_numpy(jr.split(jr.PRNGKey(0), 7).view(np.int32).copy())      # 7 seeds over 2 ranks: ragged
    local, r, w = _runner._shard_seeds(rngs)
    assert (r, w) == (rank, world)
    # every rank contributes its slice; the gathered result must be the unsharded key array
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]]))
    pad = torch.zeros((4, 2), dtype=torch.int32)
    pad[:local.shape[0]] = local
    gathered = [torch.zeros((4, 2), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, pad)
    full = torch.cat([g[:int(n)] for g, n in zip(gathered, sizes)])
    assert torch.equal(full, rngs), "seed shards do not tile the seed axis"
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([local.shape[0]]))
    dist.destroy_process_group()


def test_seed_sharding_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    n = [int(np.load(tmp_path / f"ok{r}.npy")[0]) for r in range(2)]
    assert n == [4, 3]


def _grad_worker(rank, world, port, out_dir):
    """Env-sharded data parallelism (engine.py `env_shard`): every rank computes the mean-loss gradient of ITS rows of
    a minibatch, one all-reduce(mean) of the flat gradient follows -> must equal the gradient of the union batch."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pqn_ref as R
    from purejaxql_b200 import _runner
    rng = np.random.default_rng(0)                       # same data on every rank
    B = 64
    obs = (rng.random((B, 10, 10, 4)) < 0.1).astype(np.float32)
    act = rng.integers(0, 3, B).astype(np.int32)
    tgt = rng.standard_normal(B).astype(np.float32)
    p = R.random_params(R.cnn_param_shapes(4, 3), 1)
    lo, hi = rank * B // world, (rank + 1) * B // world
    loss, _, g = R.cnn_loss_and_grads(p, obs[lo:hi], act[lo:hi], tgt[lo:hi])
    keys = sorted(g)
    flat = torch.from_numpy(np.concatenate([g[k].ravel() for k in keys]).astype(np.float32))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    _, _, gu = R.cnn_loss_and_grads(p, obs, act, tgt)
    ref = np.concatenate([gu[k].ravel() for k in keys])
    assert np.abs(flat.numpy() - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
    # mode selection of single_run
    assert _runner.pick_data_parallel({"NUM_SEEDS": 1, "NUM_ENVS": 128, "NUM_STEPS": 32, "NUM_MINIBATCHES": 32}, world) == "envs"
    assert _runner.pick_data_parallel({"NUM_SEEDS": 8, "NUM_ENVS": 128, "NUM_STEPS": 32, "NUM_MINIBATCHES": 32}, world) == "seeds"
    assert _runner.pick_data_parallel({"NUM_SEEDS": 8, "NUM_ENVS": 128, "NUM_STEPS": 32, "NUM_MINIBATCHES": 32,
                                       "DATA_PARALLEL": "envs"}, world) == "envs"
    try:
        _runner.pick_data_parallel({"NUM_SEEDS": 1, "NUM_ENVS": 127, "NUM_STEPS": 32, "NUM_MINIBATCHES": 32}, world)
        raise AssertionError("an uneven env split must be refused")
    except ValueError:
        pass
    np.save(os.path.join(out_dir, f"g{rank}.npy"), flat.numpy())
    dist.destroy_process_group()


def test_env_sharded_gradient_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.array_equal(g0, g1), "ranks must hold bit-identical averaged gradients"
