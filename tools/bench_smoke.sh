#!/bin/bash
# quick pass over every bench configuration at reduced size (syntax / plumbing check, not a measurement)
set -u
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/r2_smoke_$name.json 2> gpurun_out/r2_smoke_$name.err; echo "== $name rc=$?"; tail -c 400 gpurun_out/r2_smoke_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_smoke_$name.json"))
    print({k:d.get(k) for k in ("value","ms_per_step","stage_ms","parity")}, d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("wrapper_driven") or {}).get("value"))
    if "sweep" in d: print(d["sweep"][:2])
except Exception as e: print("no json", e)
PY
}
run C2 --config C2 --pairs 65536 --steps 2 --warmup 1 --e2e-steps 2 --cpu-sample 2048
run C3 --config C3 --pairs 4000 --steps 2 --warmup 1 --e2e-steps 2 --cpu-sample 1024
run S1 --config S1 --pairs 256 --steps 2 --warmup 1 --e2e-steps 2 --cpu-sample 64
run S2 --config S2 --pairs 200000 --steps 2 --warmup 1 --e2e-steps 2 --cpu-sample 8192
run C5 --config C5 --pairs 400000 --steps 2 --warmup 1 --e2e-steps 2 --cpu-sample 4096
run C4 --config C4 --pairs 4096 --steps 1 --warmup 1 --e2e-steps 2
