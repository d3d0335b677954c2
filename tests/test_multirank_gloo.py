"""N > 1 host logic on CPU: two gloo ranks shard a batch with shard_bounds,
"process" their slab, and reduce timing / counts exactly the way bench.py does
(max over ranks of the elapsed time, whole-job units / that time).  The data
path itself has no collective (every matrix is independent)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nitorch_fastmath_b200.shard import shard_bounds
from oracle import generators as G
from oracle import ref_port as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        mat = G.spd_packed(batch, n, torch.float64, seed=3)      # same global problem on every rank
        vec = G.vectors(batch, n, torch.float64, seed=4)
        b, e = shard_bounds(batch, world, rank, align=64)
        # stand-in for the device kernel on this rank's slab (CPU test: the oracle)
        x = P.sym_solve(mat[b:e], vec[b:e])
        elapsed = torch.tensor([1.0 + rank], dtype=torch.float64)  # rank 1 is "slower"
        dist.barrier()
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        count = torch.tensor([e - b], dtype=torch.int64)
        dist.all_reduce(count, op=dist.ReduceOp.SUM)
        gathered = [None] * world
        dist.all_gather_object(gathered, (b, e, x))
        if rank == 0:
            # plain numpy payload: torch tensors would travel by fd-passing and die with this process
            q.put((float(elapsed.item()), int(count.item()), [(b0, e0, t.numpy()) for b0, e0, t in gathered]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    world, batch, n = 2, 1000, 3
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    elapsed, count, gathered = q.get()
    assert elapsed == 2.0                     # max over ranks
    assert count == batch                     # whole-job units
    mat = G.spd_packed(batch, n, torch.float64, seed=3)
    vec = G.vectors(batch, n, torch.float64, seed=4)
    whole = P.sym_solve(mat, vec)
    stitched = torch.cat([torch.from_numpy(x) for _, _, x in sorted(gathered, key=lambda t: t[0])])
    assert torch.equal(stitched, whole)       # slabs are independent: bit-identical to one rank
    assert [g[:2] for g in sorted(gathered, key=lambda t: t[0])] == [(0, 512), (512, 1000)]
