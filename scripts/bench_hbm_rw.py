#!/usr/bin/env python
"""Pure-write / pure-read / copy HBM bandwidth of this box (measurement aid for DESIGN.md: several kernels of the
training step are write-dominated, and the driver's MEASURED_PEAKS.json number is a copy, i.e. half reads).
torch library kernels on purpose (fill_, sum, copy_): they are the yardstick, not the product."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 30                                   # 4 GiB of fp32: far beyond the 126 MB L2
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
ms = timed(lambda: a.fill_(1.0)); res["write_fill_gbs"] = 4 * n / ms / 1e6
ms = timed(lambda: a.zero_()); res["write_memset_gbs"] = 4 * n / ms / 1e6
ms = timed(lambda: b.copy_(a)); res["copy_read_plus_write_gbs"] = 8 * n / ms / 1e6
ms = timed(lambda: a.sum()); res["read_sum_gbs"] = 4 * n / ms / 1e6
res = {k: round(v, 1) for k, v in res.items()}
res["bytes"] = 4 * n
print(json.dumps(res))
