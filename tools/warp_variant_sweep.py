"""warp-per-env kernel at small batches (run on the GPU box; ST_B200_LIB selects the build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

torch.cuda.set_device(0)
os.environ["ST_B200_RAM_PATH"] = "warp"
out = []
for n in (1024, 2048, 4096, 8192, 12288, 16384, 32768):
    bench.WORKLOADS["X"] = dict(n=n, kw=dict(reward_step=True, advanced_clears=True), desc="sweep")
    r = bench.time_workload("X", 200, 5, 0, 1, None)
    out.append(f"{n}:{r['ms_per_step'] * 1e3:.2f}")
print(os.environ.get("ST_B200_LIB", "default")[-20:], "  ".join(out))
