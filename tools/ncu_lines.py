"""Per-source-line instruction and stall-sample totals from `ncu -i X.ncu-rep --page source --csv
--print-source cuda,sass` (read from stdin or a file).  Usage: ncu ... | python tools/ncu_lines.py [top_n]"""
import csv
import sys


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rows = list(csv.reader(sys.stdin))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    lines, total_i, total_s = [], 0.0, 0.0
    for r in rows:
        if len(r) > ie and r[0].isdigit():
            try:
                v, s = float(r[ie]), float(r[isamp])
            except ValueError:
                continue
            lines.append((v, s, int(r[0]), r[1][:100]))
            total_i += v
            total_s += s
    print(f"total warp-instructions {total_i:.0f}, stall samples {total_s:.0f}")
    for v, s, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f"{v / total_i * 100:5.1f}% inst {s / max(total_s, 1) * 100:5.1f}% samp  L{ln:<4d} {src}")


if __name__ == "__main__":
    main()
