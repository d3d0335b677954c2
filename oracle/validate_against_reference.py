"""Pin the oracle: compare oracle/ref_port.py with the real reference.

Run in the build container (needs /root/reference):

    python -m oracle.validate_against_reference

For every routine on the hot path it feeds identical seeded inputs to the
reference (imported through oracle/load_reference.py) and to the restatement
and reports the worst absolute difference.  The closed forms (N <= 4, dense
n <= 3) and the LAPACK-backed branches are expected to agree BIT FOR BIT,
because the restatement keeps the reference's operation order.  Exit code is
non-zero on any mismatch.  tests/test_oracle.py runs the same check when the
reference is present and otherwise relies on tests/golden/.
"""
from __future__ import annotations

import sys

import torch

from . import generators as G
from . import load_reference, ref_port as P


def _maxdiff(a, b):
    if a.shape != b.shape:
        return float("inf")
    if a.numel() == 0:
        return 0.0
    return float((a.double() - b.double()).abs().max())


def run(verbose: bool = True) -> int:
    ref_sym, ref_bat, ref_sugar = load_reference.load()
    bad = 0
    rows = []

    def check(name, got, want, exact=True, tol=0.0):
        nonlocal bad
        d = _maxdiff(got, want)
        ok = (d == 0.0) if exact else (d <= tol)
        ok = ok and got.dtype == want.dtype
        rows.append((name, d, ok))
        if not ok:
            bad += 1

    for dtype in (torch.float32, torch.float64):
        tag = "f32" if dtype == torch.float32 else "f64"
        for n in range(1, 11):
            mat = G.spd_packed((7, 33), n, dtype, seed=n)
            vec = G.vectors((7, 33), n, dtype, seed=100 + n)
            inp = G.vectors((7, 33), n, dtype, seed=200 + n)
            check(f"sym_matvec n={n} {tag}", P.sym_matvec(mat, vec), ref_sym.sym_matvec(mat, vec))
            check(f"sym_addmatvec n={n} {tag}", P.sym_addmatvec(inp, mat, vec),
                  inp + ref_sym.sym_matvec(mat, vec))
            check(f"sym_solve n={n} {tag}", P.sym_solve(mat, vec), ref_sym.sym_solve(mat, vec))
            check(f"sym_invert n={n} {tag}", P.sym_invert(mat), ref_sym.sym_invert(mat))
            check(f"sym_invert diag n={n} {tag}", P.sym_invert(mat, True), ref_sym.sym_invert(mat, True))
            check(f"sym_to_full n={n} {tag}", P.sym_to_full(mat), ref_sym.sym_to_full(mat))
            # documented regulariser semantics == reference on pre-shifted diagonal
            shifted = mat.clone()
            shifted[..., :n] += 0.25
            check(f"sym_solve diag=0.25 n={n} {tag}", P.sym_solve(mat, vec, 0.25),
                  ref_sym.sym_solve(shifted, vec))
            # broadcasting of either operand
            check(f"sym_solve bcast-vec n={n} {tag}", P.sym_solve(mat, vec[0, 0]),
                  ref_sym.sym_solve(mat, vec[0, 0]))
            check(f"sym_solve bcast-mat n={n} {tag}", P.sym_solve(mat[0, 0], vec),
                  ref_sym.sym_solve(mat[0, 0], vec))
        # eps exactly as written: only N == 2 runs in the reference
        mat = G.spd_packed(50, 2, dtype, seed=5)
        vec = G.vectors(50, 2, dtype, seed=6)
        check(f"sym_solve eps-as-written n=2 {tag}", P.sym_solve_ref_eps(mat, vec, 0.1),
              ref_sym.sym_solve(mat, vec, 0.1))
        # empty batch
        check(f"sym_solve empty {tag}", P.sym_solve(mat[:0], vec[:0]), ref_sym.sym_solve(mat[:0], vec[:0]))

        for n in range(1, 11):
            a = G.dense_shifted((5, 21), n, dtype, seed=n)
            b = G.vectors((5, 21), n, dtype, seed=50 + n)
            check(f"batchinv n={n} {tag}", P.batchinv(a), ref_bat.batchinv(a))
            check(f"batchdet n={n} {tag}", P.batchdet(a), ref_bat.batchdet(a))
            check(f"batchmatvec n={n} {tag}", P.batchmatvec(a, b), ref_bat.batchmatvec(a, b))
            check(f"solvevec lu n={n} {tag}", P.solvevec(a, b, "lu"), ref_sugar.solvevec(a, b, "lu"))
            check(f"inv lu n={n} {tag}", P.inv(a, "lu"), ref_sugar.inv(a, "lu"))
            s = G.dense_spd((5, 21), n, dtype, seed=n)
            check(f"solvevec chol n={n} {tag}", P.solvevec(s, b, "chol"), ref_sugar.solvevec(s, b, "chol"))
            check(f"inv chol n={n} {tag}", P.inv(s, "chol"), ref_sugar.inv(s, "chol"))
            rhs = G.vectors((5, 21, n), 3, dtype, seed=70 + n)
            check(f"lmdiv lu k=3 n={n} {tag}", P.lmdiv(a, rhs, "lu"), ref_sugar.lmdiv(a, rhs, "lu"))
        # the closed forms the reference only dispatches to on CUDA, run on CPU
        for n in (2, 3):
            a = G.dense_shifted((5, 21), n, dtype, seed=n)
            c = a.movedim(-1, 0).movedim(-1, 0)
            inv_ref = (ref_bat.inv2 if n == 2 else ref_bat.inv3)(c).movedim(0, -1).movedim(0, -1)
            det_ref = (ref_bat.det2 if n == 2 else ref_bat.det3)(c)
            check(f"closed_inv n={n} {tag}", P.closed_inv(a), inv_ref)
            check(f"closed_det n={n} {tag}", P.closed_det(a), det_ref)

    # "next" rows: sym_outer bit for bit; sym_matmul restated as behaviour (dense
    # congruence incl. the J H J^T quirk of the unrolled k == d <= 3 branches): to rounding
    for dtype, tol in ((torch.float32, 2e-5), (torch.float64, 1e-13)):
        tag = "f32" if dtype == torch.float32 else "f64"
        x = G.vectors((4, 9), 5, dtype, seed=3)
        check(f"sym_outer {tag}", P.sym_outer(x), ref_sym.sym_outer(x))
        for k, d in ((1, 1), (2, 2), (3, 3), (4, 4), (2, 3), (4, 2), (5, 3)):
            jac = G.vectors((4, 9, k), d, dtype, seed=10 * k + d)
            h = G.spd_packed((4, 9), k, dtype, seed=k)
            want = ref_sym.sym_matmul(jac, h)
            check(f"sym_matmul k={k} d={d} {tag}", P.sym_matmul(jac, h), want, exact=False,
                  tol=tol * float(want.abs().max()))
        # sym_det is not compared: the reference takes the order from a batch dimension
        # (_impl/sym.py:434), so it raises or is wrong unless the batch size happens to
        # equal N(N+1)/2 -- e.g. N = 3 with a batch of exactly 6:
        m3 = G.spd_packed(6, 3, dtype, seed=8)
        check(f"sym_det n=3 batch=6 {tag}", P.sym_det(m3), ref_sym.sym_det(m3), exact=False,
              tol=tol * float(P.sym_det(m3).abs().max()))

    if verbose:
        for name, d, ok in rows:
            if not ok or d != 0.0:
                print(f"{'ok ' if ok else 'BAD'} {name:40s} maxdiff={d:.3e}")
        print(f"{len(rows)} checks, {bad} mismatches "
              f"({sum(1 for r in rows if r[1] == 0.0)} bit-exact)")
    return bad


if __name__ == "__main__":
    sys.exit(1 if run() else 0)
