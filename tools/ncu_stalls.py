#!/usr/bin/env python3
"""Digest of one kernel of an .ncu-rep (needs --set full --import-source on): headline counters, stall-reason totals and
the instructions with the most stall samples.  usage: python tools/ncu_stalls.py <report.ncu-rep> [top] [kernel-name regex]"""
import csv
import io
import subprocess
import sys


def main(rep, top=30, kernel=None):
    sel = ["--kernel-name", "regex:" + kernel] if kernel else []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, r = rows[0], rows[2]
    for w in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        if w in hdr:
            print(f"{w:70s} {r[hdr.index(w)]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: 0 for s in stalls}
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            if data:
                break              # next kernel's section
            continue
        if r[ix["# Samples"]] == "# Samples":
            break
        d = {s: int(r[ix[s]] or 0) for s in stalls}
        for s in stalls:
            tot[s] += d[s]
        data.append((int(r[ix["# Samples"]] or 0), r[ix["Source"]].strip(), d, int(r[ix["Instructions Executed"]] or 0)))
    T = sum(tot.values()) or 1
    print("stall samples", T)
    for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
        print(f"  {s:26s} {v:7d} {100 * v / T:5.1f}%")
    print("top instructions by samples (index, samples, executions, SASS, main stall)")
    for i, (n, s, d, ex) in sorted(enumerate(data), key=lambda x: -x[1][0])[:top]:
        t = max(d, key=d.get)
        print(f"  {i:5d} {n:6d} {ex:9d} {s[:72]:72s} {t} {d[t]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, sys.argv[3] if len(sys.argv) > 3 else None)
