import os, sys, torch
sys.path.insert(0, "/root/repo")
import nitorch_fastmath_b200 as nfm
dev="cuda:0"
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n, B in ((3, 1 << 24), (6, 1 << 22)):
    nn = n*(n+1)//2
    mat = torch.rand(B + 1, nn, device=dev) * 0.1; mat[:, :n] += 4
    vec = torch.rand(B + 1, n, device=dev)
    by = B * (nn + 2*n) * 4
    t0 = timeit(lambda: nfm.sym_solve(mat[:B], vec[:B]))
    t1 = timeit(lambda: nfm.sym_solve(mat[1:], vec[1:]))        # storage offset: 24 B / 84 B -> not 16 B aligned
    t2 = timeit(lambda: nfm.sym_solve(mat[::2], vec[::2]))       # every second matrix
    print(f"n={n}: aligned {by/t0/1e3:7.0f} GB/s | offset-by-one {by/t1/1e3:7.0f} GB/s | every 2nd {by/2/t2/1e3:7.0f} GB/s (useful bytes)")
