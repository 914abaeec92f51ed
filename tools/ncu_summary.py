#!/usr/bin/env python3
"""Per-kernel digest of an .ncu-rep (time, instructions, issue %, pipes, top stall reasons, DRAM bytes).
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

NAMES = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        print("=====", r[ki][:80])
        for n in NAMES:
            if n in hdr:
                print(f"    {n} [{units[hdr.index(n)]}] {r[hdr.index(n)]}")
        st = sorted(((float(r[hdr.index(h)].replace(",", "") or 0), h) for h in stall), reverse=True)
        print("    stalls/issue:", ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in st[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
