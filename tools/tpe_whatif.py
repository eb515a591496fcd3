"""Timing-only what-if builds of the thread-per-env kernel (lock path / hard drop compiled out): how much of the
step do the divergent rare paths cost?  (results are WRONG by construction; never shipped)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.cuda.set_device(0)
os.environ["ST_B200_RAM_PATH"] = "thread"
for n in (65536, 262144):
    bench.WORKLOADS["X"] = dict(n=n, kw=dict(reward_step=True, advanced_clears=True), desc="x")
    r = bench.time_workload("X", 60, 5, 0, 1, None, burn_in=20)
    print(os.environ.get("ST_B200_LIB", "default")[-16:], n, f"{r['ms_per_step']*1e3:.2f} us")
