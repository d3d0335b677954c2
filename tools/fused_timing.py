#!/usr/bin/env python
"""Fused Gauss-Newton kernels against the chains they replace (device-resident, fp32):
  sym_matmul_solve(J, H, g, d)   vs   sym_solve(sym_matmul(J, H), g, d)
  sym_solve_update(x, A, v, lam, alpha, diag=d)   vs   x - alpha * sym_solve(A, v, d + lam)
GB/s = algorithmic bytes of the FUSED op / time (roofline denominator: measured copy peak)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nitorch_fastmath_b200 as nfm

dev = "cuda:0"


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


for k, d, B in ((3, 3, 1 << 24), (6, 6, 1 << 22), (6, 3, 1 << 23), (4, 4, 1 << 23), (10, 3, 1 << 22)):
    j = torch.randn(B, k, d, device=dev) * 0.3 + 2 * torch.eye(k, d, device=dev)
    h = torch.rand(B, k * (k + 1) // 2, device=dev) * 0.1
    h[:, :k] += 4
    g = torch.randn(B, d, device=dev)
    r = torch.rand(B, d, device=dev)
    out = torch.empty(B, d, device=dev)
    by = B * (k * d + k * (k + 1) // 2 + 3 * d) * 4
    t_f = timeit(lambda: nfm.sym_matmul_solve(j, h, g, r, out=out))
    t_c = timeit(lambda: nfm.sym_solve(nfm.sym_matmul(j, h), g, r, out=out))
    print(f"matmul_solve k={k} d={d}: fused {t_f*1e6:8.1f} us {by/t_f/1e9:6.0f} GB/s | chain {t_c*1e6:8.1f} us | speed-up {t_c/t_f:4.2f}x")

for n, B in ((3, 1 << 24), (6, 1 << 22), (10, 1 << 22)):
    nn = n * (n + 1) // 2
    a = torch.rand(B, nn, device=dev) * 0.1
    a[:, :n] += 4
    v, x, r = torch.randn(B, n, device=dev), torch.randn(B, n, device=dev), torch.rand(B, n, device=dev)
    out = torch.empty(B, n, device=dev)
    by = B * (nn + 4 * n) * 4
    t_f = timeit(lambda: nfm.sym_solve_update(x, a, v, 0.1, 0.5, diag=r, out=out))
    t_c = timeit(lambda: torch.sub(x, nfm.sym_solve(a, v, r + 0.1), alpha=0.5, out=out))
    print(f"solve_update(diag) n={n}: fused {t_f*1e6:8.1f} us {by/t_f/1e9:6.0f} GB/s | chain {t_c*1e6:8.1f} us | speed-up {t_c/t_f:4.2f}x")
