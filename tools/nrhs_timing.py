#!/usr/bin/env python
"""lmdiv with 1..12 right-hand sides and rmdiv, device-resident: register kernels on the TMA path
(k <= 4), the register-factorisation kernel with a run-time loop over the right-hand sides (k > 4,
rmdiv), against torch.linalg.solve on the same GPU.  GB/s = algorithmic bytes (A + B + X) / time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nitorch_fastmath_b200 as nfm

dev = "cuda:0"


def timeit(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for dt in (torch.float32, torch.float64):
    for n, k in ((3, 3), (4, 4), (6, 2), (6, 6), (6, 8), (10, 5), (10, 12)):
        B = 4 << 20 if n <= 6 else 1 << 20
        a = torch.randn(B, n, n, device=dev, dtype=dt)
        a.diagonal(0, -1, -2).add_(10)
        b = torch.randn(B, n, k, device=dev, dtype=dt)
        r = torch.randn(B, k, n, device=dev, dtype=dt)
        by = B * (n * n + 2 * n * k) * a.element_size()
        t = timeit(lambda: nfm.lmdiv(a, b))
        tr = timeit(lambda: nfm.rmdiv(r, a))
        tt = timeit(lambda: torch.linalg.solve(a, b), reps=3)
        print(f"{str(dt)[6:]} n={n:2d} k={k:2d}: lmdiv {t:9.1f} us ({by / t / 1e3:6.0f} GB/s) | rmdiv {tr:9.1f} us ({by / tr / 1e3:6.0f} GB/s)"
              f" | torch.linalg.solve {tt:10.1f} us")
