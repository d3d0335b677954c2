#!/usr/bin/env python
"""Per-launch cost of the C ABI: eager ctypes calls vs a captured CUDA graph,
for batches from 1K to 4M 3x3 fp32 solves (device-resident, rotating sets)."""
import functools
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nitorch_fastmath_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
n, nn = 3, 6
for batch in (1024, 65536, 1000000, 1 << 20, 2 << 20, 4 << 20):
    nsets = max(1, min(8, (400 << 20) // (batch * 48)))
    sets = []
    for s in range(nsets):
        mat = torch.rand(batch, nn, device=dev) * 0.1
        mat[:, :n] += 4
        sets.append((mat, torch.rand(batch, n, device=dev), torch.empty(batch, n, device=dev)))
    stream = torch.cuda.Stream(device=dev)
    calls = [functools.partial(lib.nfm_sym_solve, 0, n, 2, 0, batch, m.data_ptr(), nn, v.data_ptr(), n, None, 0,
                               o.data_ptr(), n, stream.cuda_stream) for m, v, o in sets]
    K = 400
    with torch.cuda.stream(stream):
        for i in range(20):
            calls[i % nsets]()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(K):
            calls[i % nsets]()
        e1.record(stream)
        t_issue = time.perf_counter() - t0
        stream.synchronize()
        eager_us = e0.elapsed_time(e1) * 1e3 / K
        # the same K launches as one CUDA graph
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(K):
                calls[i % nsets]()
        g.replay()
        stream.synchronize()
        e0.record(stream)
        g.replay()
        e1.record(stream)
        stream.synchronize()
        graph_us = e0.elapsed_time(e1) * 1e3 / K
    ideal = batch * 48 / 6.55e12 * 1e6
    print(f"batch {batch:8d}: eager {eager_us:7.2f} us/launch (host issue {t_issue / K * 1e6:5.2f} us)  "
          f"graph {graph_us:7.2f} us/launch   ideal@6.55TB/s {ideal:6.2f} us")
