import torch, time
x = torch.empty(604 << 20, dtype=torch.uint8).pin_memory()
y = torch.empty(201 << 20, dtype=torch.uint8).pin_memory()
dx = torch.empty_like(x, device="cuda"); dy = torch.empty_like(y, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1): dx.copy_(x, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): y.copy_(dy, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / 5
run(False)
t = run(False); print(f"H2D alone 604 MiB: {t*1e3:.2f} ms  {x.numel()/t/1e9:.1f} GB/s")
t = run(True); print(f"H2D 604 MiB + D2H 201 MiB concurrently: {t*1e3:.2f} ms")
