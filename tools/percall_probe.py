"""Latency of the per-call drop-in (Aligner.align -> ssw_init/ssw_align, a device batch of one)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
b = W.bsj_refinement_pairs(64, seed=3)
bases = np.array(list("ACGTN"))
refs = ["".join(bases[b.ref(i)]) for i in range(len(b))]
qs = ["".join(bases[b.query(i)]) for i in range(len(b))]
al = sw.Aligner(refs[0], 1, 1, 1, 1, report_cigar=True)
al.align(qs[0])
t0 = time.perf_counter()
for i in range(len(b)):
    al = sw.Aligner(refs[i], 1, 1, 1, 1, report_cigar=True)
    r = al.align(qs[i])
dt = (time.perf_counter() - t0) / len(b)
print("per-call Aligner(ref).align(query), 300-800 x 2000: %.3f ms/call" % (dt * 1e3))
j = W.junction_pairs(64, seed=4)
refs = ["".join(bases[j.ref(i)]) for i in range(len(j))]; qs = ["".join(bases[j.query(i)]) for i in range(len(j))]
t0 = time.perf_counter()
for i in range(len(j)):
    r = sw.Aligner(refs[i], 10, 4, 8, 2).align(qs[i])
print("per-call tiny pair (50 x 20): %.3f ms/call" % ((time.perf_counter() - t0) / len(j) * 1e3))
# the find_bsj shape through the drop-in: one clipped read end against a +-200 kb window (find_bsj.py:204-215)
rng = np.random.default_rng(9)
win = rng.integers(0, 4, 400000).astype(np.int8)
ref = "".join(bases[win])
al = sw.Aligner(ref, 1, 1, 1, 1)
qs = []
for k in range(8):
    m = int(rng.integers(40, 400)); st = int(rng.integers(0, 400000 - m))
    q, _ = W.noisy_channel(win[st:st + m].copy(), np.array([m]), rng)
    qs.append("".join(bases[q]))
al.align(qs[0])
t0 = time.perf_counter()
for q in qs:
    r = al.align(q)
print("per-call 40-400 nt vs 400 kb window: %.2f ms/call" % ((time.perf_counter() - t0) / len(qs) * 1e3))
