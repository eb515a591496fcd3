"""Cartesian sweep of environment knobs over bench workloads (run on the GPU box).
usage: python tools/knob_sweep.py C3:65536,C2:4096 ST_B200_TPE_EPW=8,16 ST_B200_TPE_L2=0,1 [T=1,32] ..."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
torch.cuda.set_device(0)
wls = [(w.split(":")[0], int(w.split(":")[1]) if ":" in w else None) for w in sys.argv[1].split(",")]
knobs = [(a.split("=")[0], a.split("=")[1].split(",")) for a in sys.argv[2:]]
for name, n in wls:
    kw = bench.WORKLOADS[name]["kw"]
    n = n or bench.WORKLOADS[name]["n"]
    for combo in itertools.product(*[v for _, v in knobs]):
        T = 1
        for (k, _), v in zip(knobs, combo):
            if k == "T":
                T = int(v)
            else:
                os.environ[k] = v
        bench.WORKLOADS["X"] = dict(n=n, kw=kw, desc="sweep")
        r = bench.time_workload("X", 40 if T == 1 else 6, 5, 0, 1, None, burn_in=60, T=T)
        print(f"{name} n={n:7d} " + " ".join(f"{k.replace('ST_B200_', '')}={v}" for (k, _), v in zip(knobs, combo)) +
              f": {r['ms_per_step'] * 1e3:8.2f} us  frac {r['roofline']['frac']:.3f}  {r['value'] / 1e9:.3f} G/s", flush=True)
