"""Static SASS mnemonics per kernel of the built library (cuobjdump -sass): the evidence that the stores are 16-byte
(STG.E.128), that the rgb writer and the staged ram path leave through TMA (UBLKCP), that records arrive by cp.async
(LDGSTS), and that programmatic dependent launch is compiled in (ACQBULK / PREEXIT).
Usage: python tools/sass_mnemonics.py > profiles/rN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gym_simpletetris_b200", "libsimpletetris_b200.so")
WATCH = ["STG.E.128", "STG.E", "STG.E.U8", "UBLKCP", "LDGSTS", "LDG", "LDS", "STS", "VOTE", "REDUX", "POPC", "FLO", "BREV",
         "SHFL", "ACQBULK", "PREEXIT", "BAR", "HMMA", "UTCMMA", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print("# cuobjdump -sass gym_simpletetris_b200/libsimpletetris_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a): static")
    print("# instruction mnemonics per kernel.  What to look for: STG.E.128 (16-byte observation stores), UBLKCP (cp.async.bulk TMA")
    print("# stores: rgb writer, staged ram path), LDGSTS (cp.async record fetch, thread-per-env kernel), VOTE/REDUX/POPC/FLO/BREV")
    print("# (bitboard engines), ACQBULK/PREEXIT (programmatic dependent launch), STL/LDL (spills: none), no HMMA/UTCMMA (nothing")
    print("# here is a contraction).")
    print(f"# arch: {arch}\n")
    name, counts, total = None, collections.Counter(), 0

    def flush():
        if name:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = dem.replace("st::", "").replace("(st::Params)", "").replace("(Params)", "").replace("void ", "")
            print(dem)
            print(f"  {total} instructions; " + ", ".join(f"{k} {counts[k]}" for k in WATCH if counts[k]) + "\n")

    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m and name:
            total += 1
            op = m.group(1)
            for k in WATCH:
                if op == k or (op.startswith(k + ".") and k not in ("STG.E",)) or (k == "STG.E" and op.startswith("STG.E") and not op.startswith("STG.E.128") and not op.startswith("STG.E.U8")):
                    counts[k] += 1
                    break
    flush()


if __name__ == "__main__":
    main()
