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
    print(f"   {k}: {v['ms']:.3f} ms  {v.get('achieved', 0):.1f} {v.get('unit', '')}  frac {v['frac']:.3f}")
if rf.get("embedding_path"):
    ep = rf["embedding_path"]
    print(f"   embedding path K1+K6: {ep['ms']:.3f} ms  {ep['achieved']:.0f} GB/s  frac {ep['frac']:.3f}; dominant: {rf['kernel']}")
