#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk
copy), SYNCS = mbarrier, plus HMMA (legacy mma.sync: must be 0).  Reads the in-tree libncf_b200.so with cuobjdump; no GPU.
usage: python tools/sass_evidence.py > profiles/r02_sass_hist.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "neural-collaborative-filtering-demo_b200", "libncf_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "HMMA", "LDGSTS", "ATOMS", "RED", "SHFL")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    print(f"# SASS evidence of {os.path.relpath(LIB, REPO)} (sm_100a), cuobjdump -sass; counts per kernel")
    print("# " + " ".join(f"{w:>8s}" for w in ("total",) + WATCH) + "  kernel")
    tot = collections.Counter()
    for name, c in kernels.items():
        if not any(c[w] for w in WATCH[:8]) and "tc" not in name:
            continue
        print("  " + " ".join(f"{c[w]:8d}" for w in ("_total",) + WATCH) + "  " + name)
    for c in kernels.values():
        tot.update(c)
    print("  " + " ".join(f"{tot[w]:8d}" for w in ("_total",) + WATCH) + "  ALL KERNELS OF THE LIBRARY (" + str(len(kernels)) + ")")


if __name__ == "__main__":
    main()
