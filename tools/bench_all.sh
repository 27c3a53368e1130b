#!/bin/bash
# every bench configuration at full size; JSON lines land in gpurun_out/r2_<tag>_<config>.json
set -u
tag=${1:-full}; shift || true
cfgs=${@:-C2 C3 S1 S2 C4 C5}
mkdir -p gpurun_out
for c in $cfgs; do
  timeout 1500 python bench.py --config $c > gpurun_out/r2_${tag}_$c.json 2> gpurun_out/r2_${tag}_$c.err
  echo "== $c rc=$?"; tail -c 300 gpurun_out/r2_${tag}_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_${tag}_$c.json"))
    print({k:d.get(k) for k in ("value","ms_per_step","stage_ms","pairs_per_second")}, "e2e", d["e2e"]["value"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "wrap", ((d.get("cpu_baseline") or {}).get("wrapper_driven") or {}).get("value"), "frac", d["roofline"]["frac"], d["roofline"].get("whole_step_frac"))
    for r in d.get("sweep", []): print("  ", {k:(round(v,1) if isinstance(v,float) else v) for k,v in r.items() if k!="stage_ms"})
except Exception as e: print("no json", e)
PY
done
