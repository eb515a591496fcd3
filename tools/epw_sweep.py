"""envs-per-warp sweep of the thread-per-env kernel (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.cuda.set_device(0)
os.environ["ST_B200_RAM_PATH"] = "thread"
for gname, kw, ns in (("10x20", dict(reward_step=True), (32768, 65536)), ("20x40", dict(width=20, height=40), (32768, 65536, 131072))):
    for n in ns:
        row = []
        for epw in (4, 8, 16):
            os.environ["ST_B200_TPE_EPW"] = str(epw)
            bench.WORKLOADS["X"] = dict(n=n, kw=kw, desc="x")
            r = bench.time_workload("X", 60, 5, 0, 1, None)
            row.append(f"epw{epw} {r['ms_per_step']*1e3:7.2f}")
        print(gname, n, "  ".join(row), flush=True)
