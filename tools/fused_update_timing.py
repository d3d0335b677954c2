#!/usr/bin/env python
"""Fused x -= alpha (A + lam I)^-1 v  vs the chain sym_solve -> torch update (256^3, 3x3 and 192^3, 6x6)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nitorch_fastmath_b200 as nfm
dev = "cuda:0"
def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n, side in ((3, 256), (6, 192)):
    nn, B = n * (n + 1) // 2, side ** 3
    mat = torch.rand(B, nn, device=dev) * 0.1; mat[:, :n] += 4
    vec = torch.rand(B, n, device=dev); x = torch.rand(B, n, device=dev); step = torch.empty_like(vec)
    def chain():
        nfm.sym_solve(mat, vec, 0.1, out=step)
        x.sub_(step, alpha=0.5)
    t_chain = timeit(chain)
    t_fused = timeit(lambda: nfm.sym_solve_update_(x, mat, vec, 0.1, 0.5))
    by = B * (nn + 3 * n) * 4
    print(f"n={n}: chain {t_chain:7.1f} us | fused {t_fused:7.1f} us ({by / t_fused / 1e3:6.0f} GB/s of {nn + 3 * n}*4 B/matrix) | x{t_chain / t_fused:.2f}")
