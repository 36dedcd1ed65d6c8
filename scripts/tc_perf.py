#!/usr/bin/env python
"""Time the tcgen05 3xTF32 GEMM at the bench geometry (S=128 seeds)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from purejaxql_b200 import _lib
dev = torch.device("cuda:0"); L = _lib.lib()
def run(S, M, N, K, a_mn, b_mn, split3, iters=10):
    a = torch.randn((S, K, M) if a_mn else (S, M, K), device=dev); b = torch.randn((S, K, N) if b_mn else (S, N, K), device=dev) * 0.05
    al, bl = torch.empty_like(a), torch.empty_like(b)
    L.pqn_tc_split_lo(_lib.p(a), _lib.p(al), a.numel(), None); L.pqn_tc_split_lo(_lib.p(b), _lib.p(bl), b.numel(), None)
    d = torch.empty((S, M, N), device=dev)
    f = lambda: _lib.check(L.pqn_tc_gemm_test(_lib.p(a), _lib.p(al), _lib.p(b), _lib.p(bl), _lib.p(d), S, M, N, K, a_mn, b_mn, split3, None))
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * S * M * N * K
    return {"S": S, "M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "split3": split3, "ms": round(ms, 4), "alg_tflops": round(fl / ms / 1e9, 1)}
out = []
for split3 in (1, 0):
    out.append(run(128, 4096, 128, 1024, 0, 1, split3))   # forward
    out.append(run(128, 1024, 128, 4096, 1, 1, split3))   # wgrad
    out.append(run(128, 4096, 1024, 128, 0, 0, split3))   # dgrad
print(json.dumps(out, indent=1))
