#!/bin/bash
# C5 (one 10 M-pair mixed batch in 8 slabs) on N GPUs: torchrun ranks (strong scaling, the driver's launch form) and,
# with "multi", one process driving N devices through ssw_align_batch_multi.  JSON lines land in gpurun_out/.
set -u
N=$1; mode=${2:-torchrun}
mkdir -p gpurun_out
if [ "$mode" = "multi" ]; then
  timeout 1500 python bench.py --config C5 --gpus $N --steps 2 --warmup 1 --e2e-steps 2 --no-cpu > gpurun_out/r2_c5_multi_n$N.json 2> gpurun_out/r2_c5_multi_n$N.err
  tail -c 300 gpurun_out/r2_c5_multi_n$N.err; cut -c1-400 gpurun_out/r2_c5_multi_n$N.json
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --config C5 --gpus $N --steps 3 --warmup 2 --e2e-steps 3 > gpurun_out/r2_c5_n$N.json 2> gpurun_out/r2_c5_n$N.err
  tail -c 300 gpurun_out/r2_c5_n$N.err; python - <<PY
import json
for l in open("gpurun_out/r2_c5_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N", d["value"], d["ms_per_step"], d["stage_ms"], "e2e", d["e2e"]["value"], d.get("e2e_packed",{}).get("value"))
PY
fi
