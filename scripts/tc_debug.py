#!/usr/bin/env python
"""Diagnose the tcgen05 path on a B200: dumps what TMA wrote and what one tile MMA produced."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from purejaxql_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
rng = np.random.default_rng(0)
for a_mn, b_mn in ((0, 0), (0, 1), (1, 1)):
    for nk in (1, 4):
        A = rng.integers(-4, 5, (128, 32)).astype(np.float32)      # logical [M][K], small ints => exact in tf32
        B = rng.integers(-4, 5, (32, 128)).astype(np.float32)      # logical [K][N]
        a_st = np.ascontiguousarray(A.T) if a_mn else A
        b_st = B if b_mn else np.ascontiguousarray(B.T)
        ta, tb = torch.from_numpy(a_st).to(dev), torch.from_numpy(b_st).to(dev)
        da = torch.zeros(4096, device=dev); db = torch.zeros(4096, device=dev)
        d = torch.full((128, 128), float("nan"), device=dev)
        info = torch.zeros(8, dtype=torch.int32, device=dev)
        rc = L.pqn_tc_debug(_lib.p(ta), _lib.p(tb), _lib.p(da), _lib.p(db), _lib.p(d), _lib.p(info), a_mn, b_mn, nk, None)
        torch.cuda.synchronize()
        print(f"=== a_mn={a_mn} b_mn={b_mn} nk={nk} rc={rc} info={[hex(int(x) & 0xffffffff) for x in info.cpu()]}")
        d = d.cpu().numpy(); sa = da.cpu().numpy(); sb = db.cpu().numpy()
        ref = A[:, :8 * nk] @ B[:8 * nk, :]
        print("  D == ref:", np.array_equal(d, ref), " max|D|", np.abs(d).max(), " nonzero frac", (d != 0).mean(),
              " max err", np.abs(d - ref).max())
        # smem layout check (K-major: row r at r*32 floats, 16B chunk c stored at chunk c ^ (r%8))
        def unswz_k(s):
            t = s.reshape(128, 8, 4)
            out = np.zeros_like(t)
            for r in range(128):
                for c in range(8):
                    out[r, c] = t[r, c ^ (r % 8)]
            return out.reshape(128, 32)
        def unswz_mn(s):   # 4 boxes [32 k][32 mn], 32-byte chunks XORed with k%4
            t = s.reshape(4, 32, 4, 8)
            out = np.zeros((32, 128), np.float32)
            for j in range(4):
                for k in range(32):
                    for c in range(4):
                        out[k, j * 32 + c * 8:j * 32 + c * 8 + 8] = t[j, k, c ^ (k % 4)]
            return out
        ea = np.array_equal(unswz_mn(sa), a_st) if a_mn else np.array_equal(unswz_k(sa), a_st)
        eb = np.array_equal(unswz_mn(sb), b_st) if b_mn else np.array_equal(unswz_k(sb), b_st)
        print("  smem A as expected:", ea, " smem B as expected:", eb, " |sa|max", np.abs(sa).max(), "|sb|max", np.abs(sb).max())
        if not np.array_equal(d, ref) and np.abs(d).max() > 0:
            # try to explain D by candidate reinterpretations
            for name, cand in (("A@B full K32", A @ B), ("(A@B).T", (A[:, :8 * nk] @ B[:8 * nk, :]).T)):
                print("   candidate", name, np.array_equal(d, cand))
            print("   D[0,:8]", d[0, :8], " ref[0,:8]", ref[0, :8])
