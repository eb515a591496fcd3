"""Warp-instructions and stall samples per source region of the thread-per-env kernel, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv`.  Usage: python tools/ncu_regions.py f.csv n_warps"""
import csv
import sys

RANGES = [(108, 124, "ColOps"), (126, 144, "rec col/set_col"), (155, 181, "piece_cols"), (183, 195, "cells"),
          (198, 214, "collisions"), (216, 247, "spawn"), (251, 268, "scan"), (272, 339, "lock"), (343, 388, "engine_step"),
          (390, 455, "fetch/helpers"), (460, 530, "prologue+fetch"), (531, 595, "reward/overlay"), (596, 643, "obs direct"),
          (644, 703, "obs staged"), (704, 723, "info"), (724, 748, "unset/reset"), (749, 772, "record store"),
          (773, 790, "epilogue")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    nw = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    cur, data, ie = None, {}, None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1]
        elif r and r[0] == "Line No":
            ie, isamp, ite = r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed")
        elif cur and ie and r and r[0].isdigit() and len(r) > ie:
            try:
                data.setdefault(cur, []).append((int(r[0]), float(r[ie]), float(r[isamp]), float(r[ite])))
            except ValueError:
                pass
    for f, l in data.items():
        print(f"{f}: {sum(x[1] for x in l) / nw:.1f} instr/warp, {sum(x[2] for x in l):.0f} samples")
    tpe = [k for k in data if "tpe" in k]
    if not tpe:
        return
    for a, b, name in RANGES:
        sel = [x for x in data[tpe[0]] if a <= x[0] <= b]
        v, s, t = sum(x[1] for x in sel), sum(x[2] for x in sel), sum(x[3] for x in sel)
        print(f"  {name:16s} {v / nw:8.1f} instr/warp  {s:5.0f} samples  avg active threads {t / max(v, 1):.1f}")
    for f in data:
        if f != tpe[0]:
            for x in sorted(data[f], key=lambda x: -x[1])[:10]:
                print(f"    {f.split('/')[-1]} L{x[0]}: {x[1] / nw:.1f} instr/warp, {x[2]:.0f} samples")


if __name__ == "__main__":
    main()
