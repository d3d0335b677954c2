import sys, torch
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
for dt in (torch.float32, torch.float64):
  for n, k in ((3, 3), (4, 3), (4, 4), (6, 2), (6, 6)):
    B = 4 << 20
    a = torch.randn(B, n, n, device=dev, dtype=dt); a.diagonal(0, -1, -2).add_(10)
    b = torch.randn(B, n, k, device=dev, dtype=dt)
    es = a.element_size()
    by = B * (n*n + 2*n*k) * es
    t = timeit(lambda: nfm.lmdiv(a, b))
    t1 = timeit(lambda: nfm.solvevec(a, b[..., 0].contiguous()))
    tt = timeit(lambda: torch.linalg.solve(a, b), reps=3)
    print(f"{str(dt)[6:]} n={n} k={k}: lmdiv {t:9.1f} us ({by/t/1e3:6.0f} GB/s) | single-rhs solve {t1:8.1f} us | torch.linalg.solve {tt:10.1f} us")
