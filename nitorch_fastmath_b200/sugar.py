"""Drop-in for the LU / Cholesky part of ``nitorch_fastmath.sugar``
(reference: nitorch_fastmath/sugar.py:75-137 lmdiv, :140-191 rmdiv,
:194-258 inv, :261-287 matvec, :290-341 solvevec) for batches of small square
matrices (order <= 10) on CUDA.  Only ``method='lu'`` and ``'chol'`` exist
here: the SVD / pseudo-inverse methods and non-square systems are outside
the hot path and raise ``NotImplementedError``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _dispatch as D
from . import _host, _lib
from .batched import batchinv, batchmatvec

__all__ = ['lmdiv', 'rmdiv', 'inv', 'matvec', 'solvevec']


def _algo(method: str) -> int:
    kind = method.lower()
    if kind.startswith('lu'):
        return _lib.ALGO_LU
    if kind.startswith('chol'):
        return _lib.ALGO_LDL
    if kind.startswith(('svd', 'pinv')):
        raise NotImplementedError(f"method {method!r} is outside the B200 hot path (only 'lu' and 'chol')")
    raise ValueError('Unknown inversion method {}.'.format(method))


def lmdiv(a: Tensor, b: Tensor, method: str = 'lu', rcond: float = 1e-15, out: Optional[Tensor] = None) -> Tensor:
    r"""Left matrix division ``inv(a) @ b``  (reference sugar.py:75-137).

    a : `(..., n, n)`, b : `(..., n, k)` -> `(..., n, k)`
    """
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.shape[-1] != a.shape[-2]:
        raise NotImplementedError("non-square systems (pseudo-inverse) are outside the B200 hot path")
    algo = _algo(method)
    dev = D.common_device(a, b)
    if dev.type != "cuda":
        cuda = _host.offload_device()
        n, k = a.shape[-1], b.shape[-1]
        plain = (a.dtype == b.dtype and a.dtype in D._DTYPE_CODE and a.is_contiguous() and b.is_contiguous()
                 and 1 <= n <= _lib.MAX_N and b.shape[-2] == n and a.shape[:-2] == b.shape[:-2] and k > 0 and a.numel() > 0
                 and (out is None or (out.device.type == "cpu" and out.is_contiguous() and out.dtype == a.dtype
                                      and out.shape == b.shape)))
        if plain:
            # dense host batches are streamed through the GPU in chunks (nfm_batch_solve_host)
            code = D.dtype_code(a.dtype)
            nb = a.numel() // (n * n)
            res = out if out is not None else torch.empty(b.shape, dtype=a.dtype, pin_memory=True)
            _host.run_host("nfm_batch_solve_host", code, a.dtype, n * n + n * k, n * k, nb,
                           lambda fn, ws, wsb, chunk, nbuf, streams: fn(code, n, k, algo, nb, a.data_ptr(), b.data_ptr(),
                                                                         res.data_ptr(), ws, wsb, chunk, nbuf, streams))
            return res
        r = lmdiv(a.to(cuda), b.to(cuda), method=method).cpu()
        if out is not None:
            out.copy_(r)
            return out
        return r
    n = a.shape[-1]
    if not 1 <= n <= _lib.MAX_N:
        raise ValueError(f"matrix order {n} is outside the supported range 1..{_lib.MAX_N}")
    if b.shape[-2] != n:
        raise ValueError("a and b have incompatible shapes")
    k = b.shape[-1]
    cdt = D.compute_dtype(a, b)
    batch = tuple(torch.broadcast_shapes(a.shape[:-2], b.shape[:-2]))
    nb = D.batch_count(batch)
    o, res, copy_back = D.out_operand(out, (*batch, n, k), 2, cdt, dev)
    if nb > 0 and k > 0:
        ao = D.as_operand(a, batch, 2, cdt)
        bo = D.as_operand(b, batch, 2, cdt)
        with D.device_of(dev):
            rc = _lib.load().nfm_batch_solve(D.dtype_code(cdt), n, k, algo, nb, ao.ptr, ao.stride, bo.ptr, bo.stride,
                                             o.ptr, o.stride, D.current_stream_ptr(dev))
        _lib.check(rc, "nfm_batch_solve")
    if copy_back:
        res.copy_(o.tensor)
    return res


def rmdiv(a: Tensor, b: Tensor, method: str = 'lu', rcond: float = 1e-15, out: Optional[Tensor] = None) -> Tensor:
    r"""Right matrix division ``a @ inv(b)``  (reference sugar.py:140-191, its documented meaning;
    as written the reference returns ``lmdiv(b, a)^T = (inv(b) @ a)^T`` and only runs for square ``a``).

    a : `(..., k, n)`, b : `(..., n, n)` -> `(..., k, n)`.  Solved natively as ``b^T x_r = a_r`` for every
    row ``r`` (``nfm_batch_rsolve``): neither operand is transposed or copied.
    """
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if b.shape[-1] != b.shape[-2]:
        raise NotImplementedError("non-square systems (pseudo-inverse) are outside the B200 hot path")
    algo = _algo(method)
    dev = D.common_device(a, b)
    if dev.type != "cuda":
        cuda = _host.offload_device()
        r = rmdiv(a.to(cuda), b.to(cuda), method=method).cpu()
        if out is not None:
            out.copy_(r)
            return out
        return r
    n = b.shape[-1]
    if not 1 <= n <= _lib.MAX_N:
        raise ValueError(f"matrix order {n} is outside the supported range 1..{_lib.MAX_N}")
    if a.shape[-1] != n:
        raise ValueError("a and b have incompatible shapes")
    k = a.shape[-2]
    cdt = D.compute_dtype(a, b)
    batch = tuple(torch.broadcast_shapes(a.shape[:-2], b.shape[:-2]))
    nb = D.batch_count(batch)
    o, res, copy_back = D.out_operand(out, (*batch, k, n), 2, cdt, dev)
    if nb > 0 and k > 0:
        ao = D.as_operand(a, batch, 2, cdt)
        bo = D.as_operand(b, batch, 2, cdt)
        with D.device_of(dev):
            rc = _lib.load().nfm_batch_rsolve(D.dtype_code(cdt), n, k, algo, nb, bo.ptr, bo.stride, ao.ptr, ao.stride,
                                              o.ptr, o.stride, D.current_stream_ptr(dev))
        _lib.check(rc, "nfm_batch_rsolve")
    if copy_back:
        res.copy_(o.tensor)
    return res


def inv(a: Tensor, method: str = 'lu', rcond: float = 1e-15, out: Optional[Tensor] = None) -> Tensor:
    r"""Matrix inversion (reference sugar.py:194-258).  ``'lu'`` is LAPACK
    ``getri`` in the reference, so no closed-form determinant shift here."""
    a = torch.as_tensor(a)
    if a.shape[-1] != a.shape[-2]:
        raise NotImplementedError("non-square systems (pseudo-inverse) are outside the B200 hot path")
    algo = _algo(method)
    return batchinv(a, method='lu' if algo == _lib.ALGO_LU else 'chol', regularise=False, out=out)


def matvec(mat: Tensor, vec: Tensor, out: Optional[Tensor] = None) -> Tensor:
    r"""Matrix-vector product with broadcasting (reference sugar.py:261-287)."""
    return batchmatvec(mat, vec, out=out)


def solvevec(mat: Tensor, vec: Tensor, method: str = 'lu', rcond: float = 1e-15,
             out: Optional[Tensor] = None) -> Tensor:
    r"""Left matrix-vector division ``inv(mat) @ vec``  (reference sugar.py:290-341).

    mat : `(..., n, n)`, vec : `(..., n)` -> `(..., n)`
    """
    vec = torch.as_tensor(vec).unsqueeze(-1)
    if out is not None:
        out = out.unsqueeze(-1)
    return lmdiv(mat, vec, method=method, rcond=rcond, out=out).squeeze(-1)
