"""Tiny driver for ncu: one resident batch of the C2 workload, a few runs of the hot path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kind = sys.argv[3] if len(sys.argv) > 3 else "c2"
if kind == "c2":
    b = W.bsj_refinement_pairs(n, seed=5)
elif kind == "c3":
    import torch
    b = W.repack(W.rolling_circle_pairs_torch(n, torch.device("cuda", 0), seed=5))
elif kind == "s2":
    import torch
    b = W.junction_pairs_torch(n, torch.device("cuda", 0), seed=5)
elif kind == "c5":
    import torch
    b = W.mixed_slab_torch(n, torch.device("cuda", 0), seed=5)
else:
    b = W.square_pairs(n, int(kind), params=(10, 4, 8, 2))
with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, b.match, b.mismatch, b.gap_open, b.gap_extend, flag=1) as d:
    for _ in range(runs):
        d.run()
    ms = d.stage_ms()
    rec, cig = d.fetch()
print("pairs", len(b), "cells", b.cells, "stage_ms", ms.tolist(), "GCUPS fwd", b.cells / ms[0] / 1e6, "all", b.cells / ms.sum() / 1e6,
      "bad status", int(((rec["status"] & 0xff) != 0).sum()))
