#!/usr/bin/env python
"""Throughput of the element-strided (coefficient-first, 'SoA') path vs the AoS
TMA path for the 256^3 3x3 and 192^3 6x6 solves, through the public API."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nitorch_fastmath_b200 as nfm
dev = "cuda:0"
for n, side in ((3, 256), (6, 192), (10, 160)):
    nn, B = n * (n + 1) // 2, side ** 3
    mat = torch.rand(B, nn, device=dev) * 0.1
    mat[:, :n] += 4
    vec = torch.rand(B, n, device=dev)
    mat_cf, vec_cf = mat.t().contiguous(), vec.t().contiguous()
    out_cf = torch.empty_like(vec_cf)
    out = torch.empty_like(vec)
    def timeit(f, reps=20):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    bytes_ = B * (nn + 2 * n) * 4
    t_aos = timeit(lambda: nfm.sym_solve(mat, vec, out=out))
    t_soa = timeit(lambda: nfm.sym_solve(mat_cf.t(), vec_cf.t(), out=out_cf.t()))
    t_copy = timeit(lambda: nfm.sym_solve(mat_cf.t().contiguous(), vec_cf.t().contiguous(), out=out))
    print(f"n={n}: AoS (TMA) {t_aos:7.1f} us {bytes_/t_aos/1e3:7.1f} GB/s | SoA in place {t_soa:7.1f} us {bytes_/t_soa/1e3:7.1f} GB/s | "
          f"SoA -> .contiguous() -> AoS {t_copy:7.1f} us")
