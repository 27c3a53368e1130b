"""S1 shape (find_bsj.align_clip_segments, find_bsj.py:182-233): clipped read ends of 20-500 nt against
+-200 kb genomic windows, scoring 1/1/1/1.  argv: number of pairs, 'check' to compare with the reference."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 48
check = len(sys.argv) > 2 and sys.argv[2] == "check"
rng = np.random.default_rng(11)
genome = rng.integers(0, 4, 8_000_000).astype(np.int8)
qs, rs = [], []
for k in range(n_pairs):
    n = int(rng.integers(150000, 400001))
    a = int(rng.integers(0, len(genome) - n))
    r = genome[a:a + n]
    m = int(rng.integers(20, 500))
    st = int(rng.integers(0, n - m))
    q, _ = W.noisy_channel(r[st:st + m].copy(), np.array([m]), rng, n_frac=0.01)
    qs.append(q); rs.append(r)
b = W.from_lists(qs, rs, (1, 1, 1, 1))
for flag in (1, 4):
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, flag=flag, filterd=-1 if flag == 4 else 0) as d:
        d.run(); d.run(); ms = d.stage_ms(); rec, cig = d.fetch()
    print("S1 flag %d pairs %d cells %.3e stage_ms %s  %.0f GCUPS  bad %d" % (flag, len(b), b.cells, [round(float(x), 1) for x in ms],
          b.cells / ms.sum() / 1e6, int(((rec["status"] & 0xff) != 0).sum())))
if check:
    from oracle import oracle as O
    refl = O.RefLib(); mat = O.make_mat(1, 1); bad = 0
    t0 = time.perf_counter()
    for i in range(min(len(b), 64)):
        e = refl.align(b.query(i), b.ref(i), mat, 1, 1, flag=1)
        r = rec[i]
        got = (int(r["score1"]), int(r["ref_begin1"]), int(r["ref_end1"]), int(r["read_begin1"]), int(r["read_end1"]))
        bad += got != (e["score"], e["ref_begin"], e["ref_end"], e["read_begin"], e["read_end"])
    dt = time.perf_counter() - t0
    cells = sum(len(b.query(i)) * len(b.ref(i)) for i in range(min(len(b), 64)))
    print("checked", min(len(b), 64), "bad", bad, "reference libssw one core: %.1f GCUPS" % (cells / dt / 1e9))
