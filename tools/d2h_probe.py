"""How fast can 3.3 MB (one C2 step's observations) reach pinned host memory?  (run on the GPU box)"""
import time
import torch

n = 4096 * 200
src = torch.randn(n, device="cuda")
dst = torch.empty(n).pin_memory()
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunks = [(i * n // parts, (i + 1) * n // parts) for i in range(parts)]
    def go():
        for s, (a, b) in zip(streams, chunks):
            with torch.cuda.stream(s):
                dst[a:b].copy_(src[a:b], non_blocking=True)
        for s in streams:
            s.synchronize()
    for _ in range(20):
        go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        go()
    dt = (time.perf_counter() - t0) / 200
    print(f"{parts} stream(s): {dt * 1e6:7.1f} us per 3.28 MB -> {n * 4 / dt / 1e9:5.1f} GB/s")
# kernel writing straight to mapped host memory (what the zero-copy path does), timed by events
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
