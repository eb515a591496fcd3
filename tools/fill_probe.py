"""Write-only ceiling: GB/s of a plain fill of the same size as the rgb observation block (run on the GPU box)."""
import torch
n = 131072 * 84 * 84 * 3
x = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in (("fill_(128.0)", lambda: x.fill_(128.0)), ("zero_()", lambda: x.zero_())):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10):
        fn()
    b.record(); torch.cuda.synchronize()
    print(name, f"{n * 4 * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s write-only over {n*4/1e9:.1f} GB")
y = torch.empty_like(x)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2): y.copy_(x)
torch.cuda.synchronize(); a.record()
for _ in range(5): y.copy_(x)
b.record(); torch.cuda.synchronize()
print("copy_", f"{2 * n * 4 * 5 / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s read+write")
