"""Raw pinned host -> device rate for one frame's inputs (128 MB), alone and with the 20.7 MB image download running the other way:
the ceiling of bench.py's `e2e` on the box. Usage: python tools/h2d_rate.py"""
import torch, time, os
print("affinity", len(os.sched_getaffinity(0)))
x = torch.empty(128_000_000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
y = torch.empty(20_736_000, dtype=torch.uint8).pin_memory()
dy = torch.empty_like(y, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(50):
        with torch.cuda.stream(s1):
            d.copy_(x, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                y.copy_(dy, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 50
    print("both" if both else "h2d only", round(128e6 / dt / 1e9, 2), "GB/s H2D;", round(1 / dt, 1), "frames/s bound")
