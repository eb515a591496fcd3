"""warp-per-env vs thread-per-env ram step kernel across batch sizes (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

torch.cuda.set_device(0)
kws = {"10x20": dict(reward_step=True, advanced_clears=True), "20x40": dict(width=20, height=40)}
for gname, kw in kws.items():
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536, 262144, 1048576):
        row = []
        for path in ("warp", "thread4", "thread8", "thread16", "thread32"):
            os.environ["ST_B200_RAM_PATH"] = path[:6].rstrip("0123456789")
            os.environ["ST_B200_TPE_EPW"] = path[6:] or "32"
            bench.WORKLOADS["X"] = dict(n=n, kw=kw, desc="sweep")
            r = bench.time_workload("X", 40, 5, 0, 1, None, burn_in=100)
            row.append((r["ms_per_step"] * 1e3, r["value"], r["roofline"]["frac"]))
        print(f"{gname} n={n:8d} us/step: " + "  ".join(f"{nm} {r[0]:8.2f} ({r[2]:.3f})" for nm, r in
              zip(("warp", "t4", "t8", "t16", "t32"), row)), flush=True)
