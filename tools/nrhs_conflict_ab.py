#!/usr/bin/env python
"""Staged many-right-hand-sides kernel on record lengths that bank-conflict 8..32-way: direct layout with
lane-rotated column order vs the element-major scratch (NFM_MANY_TRANSPOSE_FROM=16 default / 64 = never)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nitorch_fastmath_b200 as nfm
from nitorch_fastmath_b200 import _lib

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


print("NFM_MANY_TRANSPOSE_FROM =", os.environ.get("NFM_MANY_TRANSPOSE_FROM", "(default 16)"))
for dt in (torch.float32, torch.float64):
    for n, k in ((4, 1), (4, 8), (6, 8), (8, 5), (8, 8), (6, 6), (10, 5), (10, 12), (5, 7)):
        B = 2 << 20 if n <= 6 else 1 << 20
        a = torch.randn(B, n, n, device=dev, dtype=dt)
        a.diagonal(0, -1, -2).add_(10)
        b = torch.randn(B, n, k, device=dev, dtype=dt)
        r = torch.randn(B, k, n, device=dev, dtype=dt)
        by = B * (n * n + 2 * n * k) * a.element_size()
        t = timeit(lambda: nfm.lmdiv(a, b)) if k > 4 else float("nan")
        pl = _lib.load().nfm_last_path_was_tma()
        tr = timeit(lambda: nfm.rmdiv(r, a))
        pr = _lib.load().nfm_last_path_was_tma()
        print(f"{str(dt)[6:]} n={n:2d} k={k:2d}: lmdiv {t:9.1f} us ({by / t / 1e3:6.0f} GB/s, path {pl}) | rmdiv {tr:9.1f} us ({by / tr / 1e3:6.0f} GB/s, path {pr})")
