#!/usr/bin/env python
"""Small, fast coverage run for compute-sanitizer (memcheck / racecheck):
every kernel family once -- TMA tiles + ragged tail, segmented layout, strided
kernel, broadcast operands, warp variant, host pipeline -- checked against the
oracle.   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nitorch_fastmath_b200 as nfm                      # noqa: E402
from oracle import generators as G                       # noqa: E402
from oracle import ref_port as P                         # noqa: E402

dev = "cuda:0"
worst = 0.0


def check(got, want, tol, rec=1):
    global worst
    e = G.rel_err(got, want, rec)
    worst = max(worst, e / tol)
    assert e <= tol, (e, tol)


for dtype, tol in ((torch.float32, 1e-5), (torch.float64, 1e-12)):
    for n, batch in ((3, 2 * 1024 + 37), (6, 512 + 5), (10, 256 + 3), (4, 1000)):
        mat = G.spd_packed(batch, n, dtype, seed=n)
        vec = G.vectors(batch, n, dtype, seed=n + 1)
        reg = G.vectors(batch, n, dtype, seed=n + 2).abs()
        dm, dv = mat.to(dev), vec.to(dev)
        check(nfm.sym_solve(dm, dv), P.sym_solve(mat, vec), tol)
        check(nfm.sym_solve(dm, dv, reg.to(dev)), P.sym_solve(mat, vec, reg), tol)
        check(nfm.sym_solve(dm, dv[0]), P.sym_solve(mat, vec[0]), tol)                 # broadcast operand
        check(nfm.sym_solve(dm[::2], dv[::2]), P.sym_solve(mat[::2], vec[::2]), tol)   # strided kernel
        check(nfm.sym_solve(dm[1:], dv[1:]), P.sym_solve(mat[1:], vec[1:]), tol)       # unaligned
        check(nfm.sym_matvec(dm, dv), P.sym_matvec(mat, vec), tol)
        check(nfm.sym_invert(dm), P.sym_invert(mat), tol)
        if n > 4:
            check(nfm.sym_solve(dm, dv, method="lu"), P.sym_solve(mat, vec), tol)
            check(nfm.sym_solve(dm, dv, method="warp"), P.sym_solve(mat, vec), tol)
        v2 = dv.clone()
        nfm.sym_solve_(dm, v2)                                                          # in place
        check(v2, P.sym_solve(mat, vec), tol)
    for n, batch in ((2, 700), (4, 300), (8, 130), (5, 200)):                           # SEG and non-SEG dense
        a = G.dense_shifted(batch, n, dtype, seed=n)
        b = G.vectors(batch, n, dtype, seed=n + 1)
        check(nfm.batchinv(a.to(dev)), P.batchinv(a), tol, 2)
        check(nfm.batchdet(a.to(dev)), P.batchdet(a), tol, 0)
        check(nfm.solvevec(a.to(dev), b.to(dev)), P.solvevec(a, b), tol)
        check(nfm.batchmatvec(a.to(dev), b.to(dev)), P.batchmatvec(a, b), tol)
        rhs = G.vectors((batch, n), 3, dtype, seed=n + 2)
        check(nfm.lmdiv(a.to(dev), rhs.to(dev)), P.lmdiv(a, rhs), tol, 2)
# host pipeline
mat = G.spd_packed(50_000, 3, torch.float32, seed=9).pin_memory()
vec = G.vectors(50_000, 3, torch.float32, seed=10).pin_memory()
check(nfm.sym_solve(mat, vec), P.sym_solve(mat, vec), 1e-5)
torch.cuda.synchronize()
print(f"sanitize_case OK, worst error / tolerance = {worst:.3f}")
