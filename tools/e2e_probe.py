"""e2e (host buffers) throughput of st_host_step for each zero-copy mask (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

torch.cuda.set_device(0)
for wl in sys.argv[1:] or ["C2"]:
    for mask in (0, 1, 3, 7):
        r = bench.time_e2e(wl, 100 if wl == "C2" else 10, 5, 0, 1, None, zero_copy=mask)
        print(wl, "zero_copy mask", mask, f"{r['value']:.4g} steps/s  {r['ms_per_step']*1e3:.1f} us/step")
