#!/usr/bin/env python3
"""Hottest SASS instructions (warp stall samples) of one kernel in an .ncu-rep captured with --import-source on.
usage: python tools/sass_hot.py report.ncu-rep kernel_name [top_n]"""
import csv
import io
import subprocess
import sys


def main(path, kernel, top_n=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    body = [r for r in rows if len(r) == len(hdr) and r[si].isdigit()]
    # keep the first kernel instance only
    seen, first = set(), []
    for r in body:
        if r[0] in seen:
            break
        seen.add(r[0])
        first.append(r)
    body = first
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in body)
    inst = sum(int(r[ie]) for r in body)
    print(f"{len(body)} SASS instructions, {inst} warp instructions executed, {tot} samples")
    agg = {hdr[i]: sum(int(r[i]) for r in body) for i in stall_cols}
    print("stall totals:", ", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
    for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][si]))[:top_n]:
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{idx:5d} {r[si]:>6} {r[ie]:>8}  {r[src].strip()[:72]:72s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
