"""CPU oracle: a restatement of the reference's batched small-matrix path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nitorch_fastmath_b200/`` imports
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.  It is the
checker, never the thing shipped or measured as the product.

What is restated (reference paths are relative to ``/root/reference``):

* ``nitorch_fastmath/_impl/sym.py``     -- packed-symmetric matvec / solve /
  invert / to_full (the reference's own CPU implementation of the ``sym_*``
  names that ``nitorch_fastmath/sym.py:37`` takes from the un-vendored
  ``jitfields`` package, unpinned in ``setup.cfg:26``).
* ``nitorch_fastmath/_impl/batched.py`` -- batchinv / batchdet / batchmatvec,
  both the CPU branch (LAPACK through torch) and the closed forms the
  reference only runs on CUDA (``inv2/inv3/det2/det3``).
* ``nitorch_fastmath/sugar.py``         -- lmdiv / solvevec / inv for the
  ``'lu'`` and ``'chol'`` methods.

Parity pin: the reference holds no golden vectors for this path
(SURVEY.md section 8c).  The pin is (a) ``oracle/validate_against_reference.py``
which imports the real reference in the build container and compares it
with this file bit for bit, and (b) the fixtures in ``tests/golden/`` that
``tests/golden/make_golden.py`` produced from the real reference.

Everything computes in the input dtype with whole-batch torch ops on
coefficient-first views, exactly the way the reference does, so that timing
this file on host cores is a fair stand-in for the reference's CPU path.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Union

import torch
from torch import Tensor

__all__ = [
    "packed_order", "packed_len_to_order", "packed_index",
    "sym_to_full", "full_to_sym", "sym_matvec", "sym_addmatvec", "sym_submatvec",
    "sym_solve", "sym_solve_ref_eps", "sym_invert",
    "batchdet", "batchinv", "batchmatvec", "closed_det", "closed_inv",
    "lmdiv", "solvevec", "inv",
    "sym_outer", "sym_matmul", "sym_det",
]


# --------------------------------------------------------------------------
# packed layout helpers  (sym.py:7-14, _impl/sym.py:25-27, :37)
# --------------------------------------------------------------------------

def packed_len_to_order(length: int) -> int:
    """Matrix order N from packed length N(N+1)/2  (_impl/sym.py:37, :83, :477)."""
    return int((math.sqrt(1 + 8 * length) - 1) // 2)


def packed_index(n: int, i: int, j: int) -> int:
    """Position of a_ij in the packed vector: diagonal first, then the rows
    of the strict upper triangle (sym.py:7-14)."""
    if i == j:
        return i
    if i > j:
        i, j = j, i
    return n + i * n - (i * (i + 1)) // 2 + (j - i - 1)


def packed_order(n: int):
    """List of (i, j) in packed order."""
    order = [(i, i) for i in range(n)]
    order += [(i, j) for i in range(n) for j in range(i + 1, n)]
    return order


def sym_to_full(mat: Tensor) -> Tensor:
    """(..., NN) packed -> (..., N, N) dense   (_impl/sym.py:16-60)."""
    mat = torch.as_tensor(mat)
    n = packed_len_to_order(mat.shape[-1])
    rows = []
    for i in range(n):
        rows.append(torch.stack(
            [mat[..., packed_index(n, i, j)] for j in range(n)], dim=-1))
    return torch.stack(rows, dim=-2)


def full_to_sym(full: Tensor) -> Tensor:
    """(..., N, N) -> (..., NN) taking the upper triangle (inverse of sym_to_full)."""
    n = full.shape[-1]
    return torch.stack([full[..., i, j] for i, j in packed_order(n)], dim=-1)


# --------------------------------------------------------------------------
# matvec  (_impl/sym.py:88-172)
# --------------------------------------------------------------------------

def sym_matvec(mat: Tensor, vec: Tensor) -> Tensor:
    """y = A v, A packed symmetric  (_impl/sym.py:134-172).

    Same accumulation order as the reference: start from diag * vec, then
    walk the strict upper triangle row by row adding a_ij v_j to y_i and
    a_ij v_i to y_j (``_sym_matvecn`` :123-131; the unrolled 2/3/4 variants
    :88-119 visit the terms per output row instead -- see ``_row_order``).
    """
    n = vec.shape[-1]
    if n == 1:
        return mat * vec
    m = torch.movedim(mat, -1, 0)
    v = torch.movedim(vec, -1, 0)
    y = m[:n] * v
    if n <= 4:
        # _sym_matvec2/3/4 (:88-119): per output row, off-diagonal terms in
        # increasing column order
        for i in range(n):
            for j in range(n):
                if j != i:
                    y[i].addcmul_(m[packed_index(n, i, j)], v[j])
    else:
        c = n
        for i in range(n):
            for j in range(i + 1, n):
                y[i].addcmul_(m[c], v[j])
                y[j].addcmul_(m[c], v[i])
                c += 1
    return torch.movedim(y, 0, -1)


def sym_addmatvec(inp: Tensor, mat: Tensor, vec: Tensor) -> Tensor:
    """inp + A v.  Name only in the reference (sym.py:31); semantics per
    SURVEY.md section 8a row a2."""
    return inp + sym_matvec(mat, vec)


def sym_submatvec(inp: Tensor, mat: Tensor, vec: Tensor) -> Tensor:
    """inp - A v  (sym.py:32)."""
    return inp - sym_matvec(mat, vec)


# --------------------------------------------------------------------------
# closed-form symmetric solves, orders 2..4  (_impl/sym.py:186-324)
# d = diagonal (d[k] = a_kk), u = strict upper triangle in packed order
# --------------------------------------------------------------------------

def _sq(x):
    return x * x


def _det_s2(d, u):
    # _sym_det2 (:186-190): -(u0^2) + d0 d1
    det = _sq(u[0]).neg_()
    det.addcmul_(d[0], d[1])
    return det


def _solve_s2(d, u, v, shape):
    # _sym_solve2 (:193-200)
    det = _det_s2(d, u)
    x = v.new_empty(shape)
    x[0] = d[1] * v[0] - u[0] * v[1]
    x[1] = d[0] * v[1] - u[0] * v[0]
    x /= det
    return x


def _det_s3(d, u):
    # _sym_det3 (:203-209); u = [a01, a02, a12]
    return d.prod(0) + 2 * u.prod(0) - (
        d[0] * _sq(u[2]) + d[2] * _sq(u[0]) + d[1] * _sq(u[1]))


def _solve_s3(d, u, v, shape):
    # _sym_solve3 (:212-226): adjugate times vec, divided by det
    det = _det_s3(d, u)
    c00 = d[1] * d[2] - _sq(u[2])
    c01 = u[1] * u[2] - d[2] * u[0]
    c02 = u[0] * u[2] - d[1] * u[1]
    c11 = d[0] * d[2] - _sq(u[1])
    c12 = u[0] * u[1] - d[0] * u[2]
    c22 = d[0] * d[1] - _sq(u[0])
    x = v.new_empty(shape)
    x[0] = c00 * v[0] + c01 * v[1] + c02 * v[2]
    x[1] = c01 * v[0] + c11 * v[1] + c12 * v[2]
    x[2] = c02 * v[0] + c12 * v[1] + c22 * v[2]
    x /= det
    return x


def _det_s4(d, u):
    # _sym_det4 (:229-248); u = [a01, a02, a03, a12, a13, a23]
    a, b, c, e, f, g = u[0], u[1], u[2], u[3], u[4], u[5]
    return (d.prod(0)
            + (_sq(a * g) + _sq(b * f) + _sq(c * e))
            + - 2 * (a * b * f * g + a * c * e * g + b * c * e * f)
            + 2 * (d[0] * e * f * g + d[1] * b * c * g
                   + d[2] * a * c * f + d[3] * a * b * e)
            - (d[0] * d[1] * _sq(g) + d[0] * d[2] * _sq(f)
               + d[0] * d[3] * _sq(e) + d[1] * d[2] * _sq(c)
               + d[1] * d[3] * _sq(b) + d[2] * d[3] * _sq(a)))


def _solve_s4(d, u, v, shape):
    # _sym_solve4 (:251-324): 6 off-diagonal + 4 diagonal cofactors
    det = _det_s4(d, u)
    a, b, c, e, f, g = u[0], u[1], u[2], u[3], u[4], u[5]
    k01 = (- d[2] * d[3] * a + d[2] * c * f + d[3] * b * e
           + a * _sq(g) - b * f * g - c * e * g)
    k02 = (- d[1] * d[3] * b + d[1] * c * g + d[3] * a * e
           + b * _sq(f) - a * f * g - c * e * f)
    k03 = (- d[1] * d[2] * c + d[1] * b * g + d[2] * a * f
           + c * _sq(e) - a * e * g - b * e * f)
    k12 = (- d[0] * d[3] * e + d[0] * f * g + d[3] * a * b
           + e * _sq(c) - a * c * g - b * c * f)
    k13 = (- d[0] * d[2] * f + d[0] * e * g + d[2] * a * c
           + f * _sq(b) - a * b * g - b * c * e)
    k23 = (- d[0] * d[1] * g + d[0] * f * e + d[1] * b * c
           + g * _sq(a) - a * b * f - a * c * e)
    x = v.new_empty(shape)
    x[0] = (d[1] * d[2] * d[3] - d[1] * _sq(g) - d[2] * _sq(f)
            - d[3] * _sq(e) + 2 * e * f * g) * v[0]
    x[0] += k01 * v[1]
    x[0] += k02 * v[2]
    x[0] += k03 * v[3]
    x[1] = (d[0] * d[2] * d[3] - d[0] * _sq(g) - d[2] * _sq(c)
            - d[3] * _sq(b) + 2 * b * c * g) * v[1]
    x[1] += k01 * v[0]
    x[1] += k12 * v[2]
    x[1] += k13 * v[3]
    x[2] = (d[0] * d[1] * d[3] - d[0] * _sq(f) - d[1] * _sq(c)
            - d[3] * _sq(a) + 2 * a * c * f) * v[2]
    x[2] += k02 * v[0]
    x[2] += k12 * v[1]
    x[2] += k23 * v[3]
    x[3] = (d[0] * d[1] * d[2] - d[0] * _sq(e) - d[1] * _sq(b)
            - d[2] * _sq(a) + 2 * a * b * e) * v[3]
    x[3] += k03 * v[0]
    x[3] += k13 * v[1]
    x[3] += k23 * v[2]
    x /= det
    return x


def _solve_core(m: Tensor, v: Tensor, diag_cf: Tensor) -> Tensor:
    """m, v coefficient-first; diag_cf = the (possibly regularised) diagonal."""
    n = v.shape[0]
    shape = [n, *torch.broadcast_shapes(m.shape[1:], v.shape[1:])]
    u = m[n:]
    if n == 1:
        return v / diag_cf                                  # :384-385
    if n == 2:
        return _solve_s2(diag_cf, u, v, shape)
    if n == 3:
        return _solve_s3(diag_cf, u, v, shape)
    if n == 4:
        return _solve_s4(diag_cf, u, v, shape)
    # N > 4: expand to dense and LU-solve with partial pivoting (:392-396)
    bshape = torch.broadcast_shapes(diag_cf.shape[1:], u.shape[1:])
    packed = torch.cat([diag_cf.expand(n, *bshape), u.expand(-1, *bshape)], 0)
    full = sym_to_full(torch.movedim(packed, 0, -1))
    rhs = torch.movedim(v, 0, -1).unsqueeze(-1)
    return torch.movedim(torch.linalg.solve(full, rhs).squeeze(-1), -1, 0)


def sym_solve(mat: Tensor, vec: Tensor,
              diag: Union[None, float, Sequence[float], Tensor] = None) -> Tensor:
    """x = (A + diag(d))^-1 v   (_impl/sym.py:327-398).

    ``diag`` carries the *documented* regulariser semantics (docstring
    :356-357, padding :379-381): a float, a sequence of up to N floats
    padded with its last value, or a tensor broadcastable to ``(..., N)``;
    entry i is added to a_ii.  The reference as written (:382, ``eps[:-1]``)
    only runs for N == 2 -- that literal behaviour is ``sym_solve_ref_eps``.
    Output dtype is vec's (:196, :215, :290).
    """
    n = vec.shape[-1]
    m = torch.movedim(mat, -1, 0)
    v = torch.movedim(vec, -1, 0)
    d = m[:n]
    if diag is not None:
        if not torch.is_tensor(diag) or diag.dim() <= 1:
            e = torch.as_tensor(diag, dtype=mat.dtype).flatten()
            e = torch.cat([e, e[-1].expand(n - len(e))])
            e = e.reshape([n] + [1] * (m.dim() - 1))
        else:
            e = torch.movedim(diag.to(mat.dtype), -1, 0)
        d = d + e
    x = _solve_core(m, v, d)
    if x.dtype != vec.dtype:
        x = x.to(vec.dtype)
    return torch.movedim(x, 0, -1)


def sym_solve_ref_eps(mat: Tensor, vec: Tensor, eps) -> Tensor:
    """The reference's eps handling exactly as written (_impl/sym.py:377-382):
    ``diag + eps[:-1]`` -- runs only for N == 2, where it adds eps[0] to both
    diagonal entries; raises for every other N."""
    n = vec.shape[-1]
    m = torch.movedim(mat, -1, 0)
    v = torch.movedim(vec, -1, 0)
    e = torch.as_tensor(eps, dtype=mat.dtype).flatten()
    e = torch.cat([e, e[-1].expand(n - len(e))])
    e = e.reshape([len(e)] + [1] * (m.dim() - 1))
    d = m[:n] + e[:-1]
    return torch.movedim(_solve_core(m, v, d), 0, -1)


def sym_invert(mat: Tensor, diag: bool = False) -> Tensor:
    """A^-1 in packed order, or only its diagonal  (_impl/sym.py:455-493):
    N solves against the unit vectors, column i supplying entries (i, j>=i)."""
    mat = torch.as_tensor(mat)
    n = packed_len_to_order(mat.shape[-1])
    out = mat.new_empty([*mat.shape[:-1], n if diag else mat.shape[-1]])
    nxt = n
    for i in range(n):
        e = mat.new_zeros(n)
        e[i] = 1
        col = sym_solve(mat, e)
        out[..., i] = col[..., i]
        if not diag:
            for j in range(i + 1, n):
                out[..., nxt] = col[..., j]
                nxt += 1
    return out


# --------------------------------------------------------------------------
# "next" rows: outer product, J^T H J, determinant  (_impl/sym.py:496-670, :401-452)
# --------------------------------------------------------------------------

def sym_outer(x: Tensor) -> Tensor:
    """x x^T in packed order (_impl/sym.py:496-528, the no-grad branch)."""
    n = x.shape[-1]
    return torch.stack([x[..., i] * x[..., j] for i, j in packed_order(n)], dim=-1)


def sym_matmul(j: Tensor, h: Tensor) -> Tensor:
    """What the reference's sym_matmul returns (_impl/sym.py:637-670).

    Documented as J^T H J with j (..., k, d) and packed h (..., k(k+1)/2); the
    general branch ``jhjn`` (:596-634) computes exactly that, but the unrolled
    branches taken for k == d in {1, 2, 3} (``jhj1/2/3`` :532-593) index the
    Jacobian the other way round and return J H J^T.  Restated as behaviour
    (a dense congruence), not term by term: agreement with the reference is to
    rounding (about 1e-15 in fp64), not bit for bit."""
    k, d = j.shape[-2:]
    hf = sym_to_full(h)
    if k == d and k <= 3:
        full = j @ hf @ j.transpose(-1, -2)
    else:
        full = j.transpose(-1, -2) @ hf @ j
    return full_to_sym(full)


def sym_det(mat: Tensor) -> Tensor:
    """Determinant of a packed symmetric matrix.  The reference's sym_det
    (_impl/sym.py:401-452) takes the matrix order from a *batch* dimension
    (:434), so it raises or is wrong unless the batch size happens to equal
    N(N+1)/2; this is the documented result, ``det(sym_to_full(mat))`` -- the
    reference's own N > 4 branch (:449-450)."""
    return torch.det(sym_to_full(mat))


# --------------------------------------------------------------------------
# dense batched  (_impl/batched.py)
# --------------------------------------------------------------------------

def _det_g2(a):
    return a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]                # det2 :22-24


def _det_g3(a):
    return (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) +  # det3 :27-32
            a[0, 1] * (a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2]) +
            a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))


def _regularised(det, a):
    # inv2/inv3 (:74-76, :94-96): det += (max|a| - min|a|) * 1e-12
    mag = a.reshape((-1,) + a.shape[2:]).abs()
    spread = mag.max(dim=0).values - mag.min(dim=0).values
    return det + spread * 1e-12


def closed_det(a: Tensor) -> Tensor:
    """The closed forms the reference runs on CUDA for n <= 3
    (_impl/batched.py:55-62), evaluated on whatever device ``a`` is on."""
    n = a.shape[-1]
    c = a.movedim(-1, 0).movedim(-1, 0)
    if n == 1:
        return c[0, 0].clone()
    if n == 2:
        return _det_g2(c)
    if n == 3:
        return _det_g3(c)
    raise ValueError("closed form only for n <= 3")


def closed_inv(a: Tensor) -> Tensor:
    """inv2 / inv3 / reciprocal (_impl/batched.py:67-98, :128), including the
    ``1e-12 * range`` shift of the determinant."""
    n = a.shape[-1]
    c = a.movedim(-1, 0).movedim(-1, 0)
    if n == 1:
        return a.reciprocal()
    f = torch.empty_like(c)
    if n == 2:
        f[0, 0] = c[1, 1]
        f[1, 1] = c[0, 0]
        f[0, 1] = -c[0, 1]
        f[1, 0] = -c[1, 0]
        det = _det_g2(c)
    elif n == 3:
        for i in range(3):
            for j in range(3):
                # cofactor of a_ji, written as the cyclic 2x2 minor (:84-92)
                r0, r1 = (j + 1) % 3, (j + 2) % 3
                c0, c1 = (i + 1) % 3, (i + 2) % 3
                f[i, j] = c[r0, c0] * c[r1, c1] - c[r0, c1] * c[r1, c0]
        det = _det_g3(c)
    else:
        raise ValueError("closed form only for n <= 3")
    f /= _regularised(det, c)[None, None]
    return f.movedim(0, -1).movedim(0, -1)


def batchdet(a: Tensor) -> Tensor:
    """Reference CPU branch: ``a.det()`` (_impl/batched.py:53-54)."""
    return a.det()


def batchinv(a: Tensor) -> Tensor:
    """Reference CPU branch: ``a.inverse()`` (_impl/batched.py:119-120)."""
    return a.inverse()


def batchmatvec(mat: Tensor, vec: Tensor) -> Tensor:
    """Reference CPU branch: sugar.matvec (_impl/batched.py:175-176 ->
    sugar.py:261-287)."""
    return torch.matmul(mat, vec.unsqueeze(-1)).squeeze(-1)


# --------------------------------------------------------------------------
# sugar.py  lu / chol only
# --------------------------------------------------------------------------

def lmdiv(a: Tensor, b: Tensor, method: str = "lu") -> Tensor:
    """A^-1 B  (sugar.py:75-137; lu :125-126, chol :127-129)."""
    kind = method.lower()
    if kind.startswith("lu"):
        return torch.linalg.solve(a, b)
    if kind.startswith("chol"):
        low = torch.linalg.cholesky(a, upper=False)
        return torch.cholesky_solve(b, low, upper=False)
    raise ValueError("oracle restates only 'lu' and 'chol'")


def solvevec(mat: Tensor, vec: Tensor, method: str = "lu") -> Tensor:
    """A^-1 b for a vector right-hand side (sugar.py:290-341)."""
    return lmdiv(mat, vec.unsqueeze(-1), method).squeeze(-1)


def inv(a: Tensor, method: str = "lu") -> Tensor:
    """A^-1  (sugar.py:194-258; lu :242-243, chol :244-250)."""
    kind = method.lower()
    if kind.startswith("lu"):
        return torch.inverse(a)
    if kind.startswith("chol"):
        # (for a.dim() == 2 the reference hands ``a`` itself, not its factor,
        # to cholesky_inverse (:245-246) -- a bug outside the batched path;
        # the oracle states the documented result for every rank)
        low = torch.linalg.cholesky(a, upper=False)
        eye = torch.eye(a.shape[-2], dtype=a.dtype, device=a.device)
        return torch.cholesky_solve(eye, low, upper=False)
    raise ValueError("oracle restates only 'lu' and 'chol'")
