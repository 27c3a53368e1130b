"""Stage times of the hot path on the other BASELINE shapes (not bench lines: sanity / pathologies)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
def run(b, flag=1):
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, b.match, b.mismatch, b.gap_open, b.gap_extend, flag=flag) as d:
        d.run(); d.run(); ms = d.stage_ms(); rec, cig = d.fetch()
    print("%-28s flag %d pairs %8d cells %.3e  stage_ms %s  total %.1f ms  %.0f GCUPS  %.0f kpairs/s  bad %d" % (
        b.name, flag, len(b), b.cells, [round(float(x), 1) for x in ms], ms.sum(), b.cells / ms.sum() / 1e6, len(b) / ms.sum(), int(((rec["status"] & 0xff) != 0).sum())))
run(W.junction_pairs(1000000, seed=5))
run(W.rolling_circle_pairs(50000, seed=5))
for L, n in ((64, 200000), (256, 50000), (1024, 8000), (4096, 1000)):
    for p in ((1, 1, 1, 1), (10, 4, 8, 2)):
        run(W.square_pairs(n, L, params=p)); 
    run(W.square_pairs(n, L, params=(10, 4, 8, 2)), flag=0)
