"""Seeded synthetic inputs for the parity tests, the fixtures and the bench.

TEST INFRASTRUCTURE (see oracle/ref_port.py).  Generators follow SURVEY.md
section 8d: packed SPD ``G G^T + N I`` (condition number below ~15), dense
``randn + 10 I`` (exactly tests/test_batched.py:81-96 of the reference).
"""
from __future__ import annotations

import torch


def gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def spd_packed(batch, n: int, dtype=torch.float32, seed: int = 0) -> torch.Tensor:
    """(*batch, N(N+1)/2) packed SPD matrices, diagonal first then the rows
    of the strict upper triangle."""
    if isinstance(batch, int):
        batch = (batch,)
    g = gen(seed)
    a = torch.randn(*batch, n, n, dtype=torch.float64, generator=g)
    full = a @ a.transpose(-1, -2)
    full.diagonal(0, -1, -2).add_(n)
    iu = torch.triu_indices(n, n, 1)
    packed = torch.cat([full.diagonal(0, -1, -2), full[..., iu[0], iu[1]]], -1)
    return packed.to(dtype).contiguous()


def sym_indefinite_packed(batch, n: int, dtype=torch.float32, seed: int = 0) -> torch.Tensor:
    """Well-conditioned symmetric *indefinite* matrices: Q diag(+-[1,2]) Q^T."""
    if isinstance(batch, int):
        batch = (batch,)
    g = gen(seed)
    q, _ = torch.linalg.qr(torch.randn(*batch, n, n, dtype=torch.float64, generator=g))
    lam = 1 + torch.rand(*batch, n, dtype=torch.float64, generator=g)
    sign = torch.where(torch.arange(n) % 2 == 0, 1.0, -1.0).to(torch.float64)
    full = (q * (lam * sign)[..., None, :]) @ q.transpose(-1, -2)
    iu = torch.triu_indices(n, n, 1)
    packed = torch.cat([full.diagonal(0, -1, -2), full[..., iu[0], iu[1]]], -1)
    return packed.to(dtype).contiguous()


def vectors(batch, n: int, dtype=torch.float32, seed: int = 1) -> torch.Tensor:
    if isinstance(batch, int):
        batch = (batch,)
    return torch.randn(*batch, n, dtype=torch.float64, generator=gen(seed)).to(dtype)


def dense_shifted(batch, n: int, dtype=torch.float32, seed: int = 0, shift: float = 10.0) -> torch.Tensor:
    """randn + 10 I  (reference tests/test_batched.py:81-96)."""
    if isinstance(batch, int):
        batch = (batch,)
    a = torch.randn(*batch, n, n, dtype=torch.float64, generator=gen(seed))
    a.diagonal(0, -1, -2).add_(shift)
    return a.to(dtype)


def dense_spd(batch, n: int, dtype=torch.float32, seed: int = 0) -> torch.Tensor:
    if isinstance(batch, int):
        batch = (batch,)
    a = torch.randn(*batch, n, n, dtype=torch.float64, generator=gen(seed))
    full = a @ a.transpose(-1, -2)
    full.diagonal(0, -1, -2).add_(n)
    return full.to(dtype)


def rel_err(x: torch.Tensor, ref: torch.Tensor, rec_dims: int = 1) -> float:
    """Worst per-matrix norm-wise relative error  max_b ||x_b - r_b|| / ||r_b||
    (SURVEY.md section 7.3 'Numerics': element-wise relative error is
    meaningless for near-zero components)."""
    x = x.detach().to("cpu", torch.float64)
    ref = ref.detach().to("cpu", torch.float64)
    if x.numel() == 0:
        return 0.0
    dims = tuple(range(-rec_dims, 0)) if rec_dims else ()
    if not dims:
        return float(((x - ref).abs() / ref.abs().clamp_min(1e-300)).max())
    num = (x - ref).pow(2).sum(dims).sqrt()
    den = ref.pow(2).sum(dims).sqrt().clamp_min(1e-300)
    return float((num / den).max())
