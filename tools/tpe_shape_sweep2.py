"""direct vs staged, epw, wpc, cap at C3 (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.cuda.set_device(0)
os.environ["ST_B200_RAM_PATH"] = "thread"
for name, n in (("C3", 65536), ("C2", 4096), ("C2", 16384), ("C5b", 65536)):
    kw = bench.WORKLOADS[name]["kw"]
    for epw in (4, 8, 16, 32):
        for wpc in (1, 2, 4, 8):
            row = []
            for cap in (0, 16 // wpc if wpc <= 8 else 1, 32 // wpc, 64 // wpc):
                os.environ.update(ST_B200_TPE_EPW=str(epw), ST_B200_TPE_WPC=str(wpc), ST_B200_TPE_CTAS_PER_SM=str(cap),
                                  ST_B200_TPE_STAGED="0")
                bench.WORKLOADS["X"] = dict(n=n, kw=kw, desc="sweep")
                r = bench.time_workload("X", 40, 5, 0, 1, None, burn_in=60)
                row.append(f"cap{cap:2d} {r['ms_per_step'] * 1e3:7.2f} ({r['roofline']['frac']:.3f})")
            print(f"{name} n={n:7d} direct epw={epw:2d} wpc={wpc}: " + "  ".join(row), flush=True)
