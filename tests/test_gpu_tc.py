"""GPU tests of the tcgen05 / TMEM / TMA GEMM path (3xTF32) against fp64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _run(S, M, N, K, a_mn, b_mn, split3, seed=0):
    from purejaxql_b200 import _lib
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((S, M, K)).astype(np.float32)          # logical [M,K]
    B = (rng.standard_normal((S, K, N)) * 0.05).astype(np.float32)  # logical [K,N]
    a_store = np.ascontiguousarray(A.transpose(0, 2, 1)) if a_mn else A
    b_store = B if b_mn else np.ascontiguousarray(B.transpose(0, 2, 1))
    ta, tb = torch.from_numpy(a_store).to(dev()), torch.from_numpy(b_store).to(dev())
    tal, tbl = torch.empty_like(ta), torch.empty_like(tb)
    L = _lib.lib()
    _lib.check(L.pqn_tc_split_lo(_lib.p(ta), _lib.p(tal), ta.numel(), _lib.stream_ptr()))
    _lib.check(L.pqn_tc_split_lo(_lib.p(tb), _lib.p(tbl), tb.numel(), _lib.stream_ptr()))
    d = torch.full((S, M, N), float("nan"), device=dev())
    _lib.check(L.pqn_tc_gemm_test(_lib.p(ta), _lib.p(tal), _lib.p(tb), _lib.p(tbl), _lib.p(d), S, M, N, K, a_mn, b_mn,
                                  split3, _lib.stream_ptr()), "pqn_tc_gemm_test")
    torch.cuda.synchronize()
    ref = np.matmul(A.astype(np.float64), B.astype(np.float64))
    return d.cpu().numpy(), ref, A, B, tal.cpu().numpy()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (1, 1), (0, 0)])
def test_tc_gemm_single_pass_tf32(a_mn, b_mn):
    d, ref, A, B, _ = _run(2, 256, 128, 96, a_mn, b_mn, 0)
    # single TF32 pass == exact product of the truncated operands (fp32 accumulate)
    trunc = lambda x: (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    ref_t = np.matmul(trunc(A).astype(np.float64), trunc(B).astype(np.float64))
    assert np.isfinite(d).all()
    assert np.abs(d - ref_t).max() < 5e-5, np.abs(d - ref_t).max()
    assert np.abs(d - ref).max() < 2e-2


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (1, 1), (0, 0)])
@pytest.mark.parametrize("S,M,N,K", [(1, 128, 128, 32), (3, 200, 128, 1024), (2, 1024, 256, 4096)])
def test_tc_gemm_inline_a_lo_is_bitwise_the_same(a_mn, b_mn, S, M, N, K):
    """split3=2 derives A_lo inside the kernel (converter warps, smem -> smem) instead of reading it from memory:
    same operands, same MMA order, so the result must be identical to the precomputed-lo run."""
    d1, ref, *_ = _run(S, M, N, K, a_mn, b_mn, 1, seed=K)
    d2, *_ = _run(S, M, N, K, a_mn, b_mn, 2, seed=K)
    assert np.isfinite(d2).all()
    assert np.array_equal(d1, d2)
    assert np.abs(d2 - ref).max() < 4e-6 * np.abs(ref).max()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (1, 1), (0, 0)])
@pytest.mark.parametrize("S,M,N,K", [(1, 128, 128, 32), (3, 200, 128, 1024), (2, 1024, 256, 4096)])
def test_tc_gemm_3xtf32_fp32_accuracy(a_mn, b_mn, S, M, N, K):
    d, ref, A, B, alo = _run(S, M, N, K, a_mn, b_mn, 1, seed=K)
    assert np.isfinite(d).all()
    scale = np.abs(ref).max()
    err = np.abs(d - ref).max()
    assert err < 4e-6 * scale, (err, scale)
    # the lo operand is what the header says
    hi = (A.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo_ref = A - hi
    lo_ref = np.ascontiguousarray(lo_ref.transpose(0, 2, 1)) if a_mn else lo_ref
    assert np.array_equal(alo, lo_ref)
