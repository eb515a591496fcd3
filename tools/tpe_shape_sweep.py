"""Launch-shape sweep of the thread-per-env kernel: envs per group x warps per CTA x grid cap (run on the GPU box).
usage: python tools/tpe_shape_sweep.py [10x20|20x40|C3] ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.cuda.set_device(0)
os.environ["ST_B200_RAM_PATH"] = "thread"
GEOMS = {"10x20": (dict(reward_step=True, advanced_clears=True), 10, 20, (4096, 16384, 65536, 262144)),
         "C3": (bench.WORKLOADS["C3"]["kw"], 10, 20, (65536,)),
         "20x40": (dict(width=20, height=40), 20, 40, (16384, 65536, 262144))}


def smem(W, H, epw, wpc):
    cw = 1 if H <= 31 else 2
    pitch = (15 + W * cw) | 1
    rec = (epw * pitch + 3) & ~3
    stage = (epw * W * H * 4 + 15) >> 4 << 2
    return wpc * (2 * rec + stage) * 4


for gname in (sys.argv[1:] or list(GEOMS)):
    kw, W, H, ns = GEOMS[gname]
    for n in ns:
        for epw in (4, 8, 16, 32):
            for wpc in (1, 2, 4, 8):
                sm = smem(W, H, epw, wpc)
                if sm > 226 * 1024:
                    continue
                occ = min(32, (227 * 1024) // (sm + 1024 + 512), 64 // wpc)
                row = []
                for cap in (0, occ, max(1, occ // 2)):
                    os.environ["ST_B200_TPE_EPW"] = str(epw)
                    os.environ["ST_B200_TPE_WPC"] = str(wpc)
                    os.environ["ST_B200_TPE_CTAS_PER_SM"] = str(cap)
                    bench.WORKLOADS["X"] = dict(n=n, kw=kw, desc="sweep")
                    r = bench.time_workload("X", 40, 5, 0, 1, None, burn_in=60)
                    row.append(f"cap{cap:2d} {r['ms_per_step'] * 1e3:8.2f} ({r['roofline']['frac']:.3f})")
                print(f"{gname} n={n:7d} epw={epw:2d} wpc={wpc} smem={sm // 1024:3d}K occ={occ:2d}: " + "  ".join(row), flush=True)
