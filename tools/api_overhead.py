#!/usr/bin/env python
"""Host-side cost of one public-API call (tiny batch, so the GPU is idle)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nitorch_fastmath_b200 as nfm
dev = "cuda:0"
mat = torch.rand(1024, 6, device=dev) + 3
vec = torch.rand(1024, 3, device=dev)
out = torch.empty_like(vec)
a = torch.rand(1024, 4, 4, device=dev, dtype=torch.float64) + 4 * torch.eye(4, device=dev, dtype=torch.float64)
def bench(name, f, n=3000):
    for _ in range(100): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    print(f"{name:40s} {(time.perf_counter() - t0) / n * 1e6:7.1f} us / call")
bench("sym_solve(mat, vec)", lambda: nfm.sym_solve(mat, vec))
bench("sym_solve(mat, vec, out=out)", lambda: nfm.sym_solve(mat, vec, out=out))
bench("sym_solve(mat, vec, 0.1, out=out)", lambda: nfm.sym_solve(mat, vec, 0.1, out=out))
bench("sym_matvec(mat, vec, out=out)", lambda: nfm.sym_matvec(mat, vec, out=out))
bench("sym_invert(mat)", lambda: nfm.sym_invert(mat))
bench("batchinv(a)", lambda: nfm.batchinv(a))
bench("torch: mat * 2 (reference point)", lambda: mat * 2)
