"""Dynamic instruction statistics of one kernel from an ncu report (SASS page).
Usage: python tools/sass_hist.py report.ncu-rep [dump_hot_threshold]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hi[0]]
ie = hdr.index("Instructions Executed")
body = [r for r in rows[hi[0] + 1:(hi[1] if len(hi) > 1 else len(rows))] if len(r) > ie and r[ie].isdigit()]
print(rows[0][:2])
hist, ops = collections.Counter(), collections.Counter()
for r in body:
    v = int(r[ie])
    hist[v] += 1
    f = r[1].split()
    ops[(f[1] if f[0].startswith("@") else f[0]).split(".")[0]] += v
tot = sum(k * v for k, v in hist.items())
base = max(hist, key=lambda k: hist[k] * k)
print(f"dynamic warp-instructions {tot}; dominant execution count {base} -> {tot / base:.1f} per env-step")
for k, v in sorted(hist.items(), key=lambda kv: -kv[0] * kv[1])[:8]:
    print(f"  executed {k:8d} times: {v:4d} static instructions = {k * v / tot * 100:5.1f}%")
print(sorted(ops.items(), key=lambda kv: -kv[1])[:20])
if len(sys.argv) > 2:
    thr = int(sys.argv[2])
    for r in body:
        if int(r[ie]) >= thr:
            print(r[ie].rjust(8), r[1].strip()[:100])
