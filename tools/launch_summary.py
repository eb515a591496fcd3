"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python tools/launch_summary.py gpurun_out/launches.csv "<command line>" > profiles/rN_launch_list_summary.csv"""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
        name = re.sub(r"\(st::Params\)|void |st::", "", r[kn])
        name = re.sub(r"at::native::.*?<(.*?)[,>].*", r"torch:\1", name)[:110]
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: {sys.argv[2] if len(sys.argv) > 2 else '?'}")
    print("# per-launch times are cold-cache and serialised: compare SHARES.  Inside the timed regions the ONLY kernel is the step"
          " kernel (one launch per step, replayed from CUDA graphs); long st_step_tpe_kernel / st_main_kernel<..,(bool)1> launches are"
          " the st_step_many burn-ins of the set-up phase (untimed) and the C2_T32 / C3_T32 modes; fills / random_ kernels are torch set-up.")
    print(f"# total device time {tot / 1e3:.2f} ms over {n} launches")
    print("kernel,launches,sum_us,avg_us,min_us,share")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"\"{k}\",{len(v)},{sum(v):.1f},{sum(v) / len(v):.2f},{min(v):.2f},{sum(v) / tot:.4f}")


if __name__ == "__main__":
    main()
