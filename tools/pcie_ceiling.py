#!/usr/bin/env python
"""What the HOST can move: plain pinned cudaMemcpyAsync in both directions at once, on
1..N GPUs concurrently (one process per GPU, as bench.py runs).  This is the ceiling of
the end-to-end (host-buffer) path: bench.py's e2e line for the 3x3 solve moves 36 B up and
12 B down per matrix.

    python tools/pcie_ceiling.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

Prints one JSON line (rank 0): aggregate and per-rank GB/s, and the matrices/s of the 3x3
solve that bandwidth allows."""
import json
import os
import time

import torch

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

UP, DOWN = 36 * (1 << 24) // world, 12 * (1 << 24) // world     # this rank's slab of the 256^3 3x3 solve
x = torch.empty(UP, dtype=torch.uint8).pin_memory()
y = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
x.fill_(1)
dx = torch.empty(UP, dtype=torch.uint8, device="cuda")
dy = torch.zeros(DOWN, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
CH = 16 << 20


def run(up, down, reps=8):
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                for o in range(0, UP, CH):
                    dx[o:o + CH].copy_(x[o:o + CH], non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                for o in range(0, DOWN, CH):
                    y[o:o + CH].copy_(dy[o:o + CH], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if dist is not None:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


run(True, True, 2)
t_up, t_down, t_both = run(True, False), run(False, True), run(True, True)
if rank == 0:
    print(json.dumps({
        "n_gpus": world, "h2d_bytes_per_step": UP * world, "d2h_bytes_per_step": DOWN * world,
        "h2d_alone_GBs": UP * world / t_up / 1e9, "d2h_alone_GBs": DOWN * world / t_down / 1e9,
        "both_ms": t_both * 1e3, "both_GBs": (UP + DOWN) * world / t_both / 1e9,
        "ceiling_matrices_per_s_3x3_solve": (1 << 24) / t_both,
        "how": "pinned host buffers, 16 MiB cudaMemcpyAsync chunks, H2D and D2H on two streams, max over ranks"}))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
