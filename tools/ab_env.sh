#!/bin/bash
# A/B loop over one environment switch of the library: tools/ab_env.sh VAR "v1 v2 ..." [bench args]; prints ms/step per value
var=$1; vals=$2; shift 2
mkdir -p gpurun_out
for v in $vals; do
  env $var=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --min-seconds 0.4 "$@" > gpurun_out/ab_${var}_$v.json 2> gpurun_out/ab_${var}_$v.err
  python - "$var" "$v" <<'P'
import json,sys
var,v=sys.argv[1:3]
try:
    t=[l for l in open(f"gpurun_out/ab_{var}_{v}.json") if l.startswith("{")]
    d=json.loads(t[-1]); print(f"{var}={v}: {d['ms_per_step']:.4f} ms/step  e2e {d['e2e']['ms_per_step']:.4f}  loss {d.get('last_loss')}")
except Exception as e:
    print(f"{var}={v}: FAILED {e}"); print(open(f"gpurun_out/ab_{var}_{v}.err").read()[-1500:])
P
done
