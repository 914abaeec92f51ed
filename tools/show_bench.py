#!/usr/bin/env python3
"""Print the essentials of a bench.py JSON line read from a file argument or stdin (or the error text)."""
import json
import sys

txt = (open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()).strip().splitlines()
line = next((ln for ln in reversed(txt) if ln.startswith("{")), None)
if line is None:
    print("NO JSON LINE:\n" + "\n".join(txt[-15:]))
    sys.exit(0)
d = json.loads(line)
print(f"value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.4g}  dtype {d['dtype']}  launches {d.get('gpu_launches')}")
rf = d.get("roofline") or {}
for k, v in (rf.get("kernels") or {}).items():
    hv = v.get("hbm_view")
    extra = f"   | hbm view {hv['achieved']:.0f} GB/s frac {hv['frac']:.3f}" if hv else ""
    print(f"   {k}: {v['ms']:.3f} ms  {v.get('achieved', 0):.1f} {v.get('unit', '')}  frac {v['frac']:.3f}{extra}")
if rf.get("embedding_path"):
    ep = rf["embedding_path"]
    print(f"   embedding path K1+K6: {ep['ms']:.3f} ms  {ep['achieved']:.0f} GB/s  frac {ep['frac']:.3f}; dominant: {rf['kernel']}")
if rf.get("whole_step_hbm_view"):
    ws = rf["whole_step_hbm_view"]
    print(f"   whole step: {ws['algorithmic_bytes'] / 1e6:.0f} MB algorithmic, {ws['achieved']:.0f} GB/s, frac {ws['frac']:.3f}; pieces sum {rf.get('pieces_sum_ms', 0):.3f} ms")
