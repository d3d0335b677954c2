#!/usr/bin/env python
"""One lmdiv shape, a few launches (for ncu): python tools/nrhs_one.py n k f32|f64 [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nitorch_fastmath_b200 as nfm

n, k = int(sys.argv[1]), int(sys.argv[2])
dt = torch.float32 if sys.argv[3] == "f32" else torch.float64
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2 << 20
a = torch.randn(B, n, n, device="cuda", dtype=dt)
a.diagonal(0, -1, -2).add_(10)
b = torch.randn(B, n, k, device="cuda", dtype=dt)
out = torch.empty_like(b)
for _ in range(4):
    nfm.lmdiv(a, b, out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0]))
