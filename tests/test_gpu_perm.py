"""GPU parity of ``pqn_permutation`` (csrc/pqn_perm.cu) with the oracle's restatement of ``jax.random.permutation``
(rounds of a stable sort by fresh 32-bit keys; purejaxql/pqn_minatar.py:299-321): bit-exact index permutations for
both threefry layouts, ragged sizes, the minibatch output layout, the oversized-bucket fallback and the full
BASELINE size (property checks)."""
import numpy as np
import pytest
import torch

from oracle import jax_prng as jr

pytestmark = pytest.mark.gpu


def tkeys(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32).copy()).to("cuda:0")


@pytest.mark.parametrize("part", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 37, 64, 65, 1000, 4096, 20000])
def test_permutation_matches_oracle_bit_exact(n, part):
    from purejaxql_b200 import jaxrandom
    keys = np.stack([jr.PRNGKey(s) for s in (0, 5, 77)])
    got = jaxrandom.permutation_indices(tkeys(keys), n, part).cpu().numpy()
    assert got.shape == (3, n) and got.dtype == np.int32
    jr.DEFAULT_PARTITIONABLE = bool(part)
    try:
        for i in range(3):
            assert np.array_equal(got[i], jr.permutation_indices(keys[i], n)), (n, part, i)
    finally:
        jr.DEFAULT_PARTITIONABLE = False


def test_permutation_minibatch_layout_and_workspace_reuse():
    from purejaxql_b200 import jaxrandom
    keys = np.stack([jr.PRNGKey(s) for s in (3, 4)])
    n, chunk = 2048, 256
    ws = jaxrandom.permutation_workspace(n, 2, "cuda:0")
    plain = jaxrandom.permutation_indices(tkeys(keys), n, 0, workspace=ws)
    mb = jaxrandom.permutation_indices(tkeys(keys), n, 0, chunk=chunk, workspace=ws)
    assert mb.shape == (n // chunk, 2, chunk)
    assert torch.equal(mb, plain.view(2, n // chunk, chunk).transpose(0, 1).contiguous())
    assert torch.equal(plain, jaxrandom.permutation_indices(tkeys(keys), n, 0, workspace=ws))   # scratch state is reset


def test_permutation_oversized_bucket_path():
    """Buckets of ~1024 elements (> the 256 a warp stages in shared memory) take the global-memory rank path."""
    from purejaxql_b200 import _lib, jaxrandom
    keys = np.stack([jr.PRNGKey(11)])
    prev = _lib.lib().pqn_set_permutation_bucket_log2(10)
    try:
        got = jaxrandom.permutation_indices(tkeys(keys), 8192, 0).cpu().numpy()
    finally:
        _lib.lib().pqn_set_permutation_bucket_log2(prev)
    assert np.array_equal(got[0], jr.permutation_indices(keys[0], 8192))


def test_permutation_full_size_properties():
    """BASELINE geometry (128 seeds x 131,072 samples): every row is a permutation, rows differ, seed 0 and seed 127
    equal the oracle."""
    from purejaxql_b200 import jaxrandom
    S, n = 128, 131072
    keys = jr.split(jr.PRNGKey(9), S)
    got = jaxrandom.permutation_indices(tkeys(keys), n, 0)
    srt = torch.sort(got, dim=1).values
    assert torch.equal(srt, torch.arange(n, device=got.device, dtype=torch.int32).expand(S, n))
    assert not torch.equal(got[0], got[1])
    g = got.cpu().numpy()
    for s in (0, 127):
        assert np.array_equal(g[s], jr.permutation_indices(keys[s], n))
