"""Warp-stall breakdown (PC sampling) and occupancy figures of every kernel in a .ncu-rep.
Usage: python tools/ncu_stalls.py gpurun_out/prof_C3.ncu-rep"""
import csv
import io
import subprocess
import sys

EXTRA = ["smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_active.avg.per_cycle_active",
         "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.avg.per_cycle_active",
         "launch__waves_per_multiprocessor", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}")
        st = []
        for i, k in enumerate(hdr):
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    st.append((float(r[i]), k[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print(f"warp stall samples (smsp__pcsamp_warps_issue_stalled_*), {tot:.0f} in total:")
        for v, k in sorted(st, reverse=True):
            if v:
                print(f"  {k:28s} {v:8.0f}  {v / tot * 100:5.1f} %")
        for k in EXTRA:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:70s} {r[i]:>16s} {units[i]}")
        print()


if __name__ == "__main__":
    main()
