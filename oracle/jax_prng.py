"""NumPy restatement of the slice of ``jax.random`` that the PQN hot path uses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED against a
live JAX (none is installable here); pinned instead against the Random123
Threefry-2x32-20 KATs and documented ``PRNGKey(0)`` outputs in
``tests/test_oracle_prng.py``.

Restated third-party algorithm: ``jax==0.4.x`` (reference pin
``pyproject.toml:27``: ``jax>=0.4.16,<=0.4.38``), ``jax/_src/prng.py`` and
``jax/_src/random.py``:

* ``threefry2x32``            — Threefry-2x32, 20 rounds (Salmon et al. 2011).
* ``split`` / ``random_bits`` — both counter layouts:
    - ``partitionable=False`` ("original", the default for every jax in the
      reference's pin range): counters ``iota(n)`` are padded to even length,
      split in halves (x0 = first half, x1 = second half), outputs are
      concatenated ``[y0..., y1...]``.
    - ``partitionable=True`` (default from jax 0.5.0, outside the pin):
      element ``i`` uses the 64-bit counter ``(hi, lo) = (0, i)``; ``split``
      keeps both output words, 32-bit ``random_bits`` xors them.
* ``uniform``, ``randint``, ``choice`` (no ``p``), ``permutation`` (``_shuffle``:
  rounds of stable sort by fresh 32-bit keys).

Reference call sites that consume these: ``purejaxql/pqn_minatar.py:116-125``
(ε-greedy split/uniform/randint), ``:107-112`` (per-env key split), ``:183``
(3-way split), ``:303`` (permutation), ``:456-459`` (PRNGKey/split over seeds).

Keys are ``uint32[..., 2]`` arrays.  All functions are vectorised over leading
key axes so a whole ``[S, E]`` batch of per-env keys is handled at once.
"""
from __future__ import annotations

import math

import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)

# Module-level default for the counter layout; the reference's pinned jax
# (<=0.4.38) uses the original (non-partitionable) layout.
DEFAULT_PARTITIONABLE = False


def _rotl(x, r):
    return ((x << U32(r)) | (x >> U32(32 - r))).astype(U32)


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 block function on broadcastable uint32 arrays."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32)
        x1 = np.asarray(x1, dtype=U32)
        ks = (k0, k1, (k0 ^ k1 ^ _PARITY).astype(U32))
        x0 = (x0 + ks[0]).astype(U32)
        x1 = (x1 + ks[1]).astype(U32)
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, r)
                x1 = (x1 ^ x0).astype(U32)
            x0 = (x0 + ks[(g + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(g + 2) % 3] + U32(g + 1)).astype(U32)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey`` for a Python int seed (x64 disabled: the high word
    is 0 for seeds < 2**32)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _bits_flat(key, n, partitionable):
    """``random_bits(key, 32, (n,))`` for keys of shape [..., 2] -> [..., n]."""
    key = np.asarray(key, dtype=U32)
    k0 = key[..., 0:1]
    k1 = key[..., 1:2]
    if partitionable:
        lo = np.arange(n, dtype=U32)
        y0, y1 = threefry2x32(k0, k1, np.zeros_like(lo), lo)
        return (y0 ^ y1).astype(U32)
    half = (n + 1) // 2
    c0 = np.arange(half, dtype=U32)
    c1 = (np.arange(half, dtype=np.uint64) + np.uint64(half)).astype(U32)
    if n % 2:
        c1 = c1.copy()
        c1[-1] = 0  # the padding counter is a literal 0
    y0, y1 = threefry2x32(k0, k1, c0, c1)
    out = np.concatenate([y0, y1], axis=-1)
    return out[..., :n]


def random_bits(key, shape=(), partitionable=None):
    if partitionable is None:
        partitionable = DEFAULT_PARTITIONABLE
    shape = tuple(shape)
    n = int(math.prod(shape)) if shape else 1
    key = np.asarray(key, dtype=U32)
    flat = _bits_flat(key, n, partitionable)
    return flat.reshape(key.shape[:-1] + shape)


def split(key, num: int = 2, partitionable=None):
    """``jax.random.split(key, num)`` -> keys of shape [..., num, 2]."""
    if partitionable is None:
        partitionable = DEFAULT_PARTITIONABLE
    key = np.asarray(key, dtype=U32)
    k0 = key[..., 0:1]
    k1 = key[..., 1:2]
    if partitionable:
        lo = np.arange(num, dtype=U32)
        y0, y1 = threefry2x32(k0, k1, np.zeros_like(lo), lo)
        return np.stack([y0, y1], axis=-1).astype(U32)
    c0 = np.arange(num, dtype=U32)
    c1 = (np.arange(num, dtype=np.uint64) + np.uint64(num)).astype(U32)
    y0, y1 = threefry2x32(k0, k1, c0, c1)
    out = np.concatenate([y0, y1], axis=-1)  # [..., 2*num]
    return out.reshape(key.shape[:-1] + (num, 2))


def uniform(key, shape=(), minval=0.0, maxval=1.0, partitionable=None):
    """``jax.random.uniform`` float32."""
    bits = random_bits(key, shape, partitionable)
    fbits = ((bits >> U32(9)) | U32(0x3F800000)).astype(U32)
    floats = fbits.view(np.float32) - np.float32(1.0)
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    out = floats * np.float32(maxval - minval) + minval
    return np.maximum(minval, out).astype(np.float32)


def randint(key, shape, minval: int, maxval: int, partitionable=None):
    """``jax.random.randint`` int32, scalar python bounds."""
    ks = split(key, 2, partitionable)
    hi = random_bits(ks[..., 0, :], shape, partitionable).astype(np.uint64)
    lo = random_bits(ks[..., 1, :], shape, partitionable).astype(np.uint64)
    span = int(maxval) - int(minval)
    if span <= 0:
        span = 1
    span &= 0xFFFFFFFF
    mult = (1 << 16) % span
    mult = (mult * mult) % span
    s = np.uint64(span)
    # all arithmetic is uint32 in jax: (hi % span) * mult wraps mod 2**32
    off = ((hi % s) * np.uint64(mult)) & np.uint64(0xFFFFFFFF)
    off = (off + (lo % s)) & np.uint64(0xFFFFFFFF)
    off = off % s
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def choice_index(key, n: int, shape=(), partitionable=None):
    """``jax.random.choice(key, a, shape)`` without ``p`` == ``a[randint(key, shape, 0, n)]``."""
    return randint(key, shape, 0, n, partitionable)


def bernoulli(key, p, shape=(), partitionable=None):
    return uniform(key, shape, partitionable=partitionable) < np.float32(p)


def permutation_indices(key, n: int, partitionable=None):
    """Index permutation produced by ``jax.random.permutation(key, x)`` on an
    array with ``x.shape[0] == n`` (``_shuffle``: ``ceil(3 ln n / ln(2**32-1))``
    rounds of *stable* sort by fresh uint32 keys).  Single key (shape [2])."""
    key = np.asarray(key, dtype=U32)
    assert key.shape == (2,)
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))
    idx = np.arange(n, dtype=np.int64)
    for _ in range(rounds):
        ks = split(key, 2, partitionable)
        key, sub = ks[0], ks[1]
        sort_keys = random_bits(sub, (n,), partitionable)
        order = np.argsort(sort_keys, kind="stable")
        idx = idx[order]
    return idx
