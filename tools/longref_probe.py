"""Forward-pass time per cell as a function of the reference length (query 300 nt, 1/1/1/1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
rng = np.random.default_rng(3)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for n, pairs in ((2000, 9472), (16000, 4736), (32000, 4736), (40000, 4736), (100000, 2368), (400000, 2368)):
    genome = rng.integers(0, 4, n + 5000).astype(np.int8)
    qs, rs = [], []
    for k in range(pairs):
        a = int(rng.integers(0, 5000))
        r = genome[a:a + n]
        st = int(rng.integers(0, n - m))
        q, _ = W.noisy_channel(r[st:st + m].copy(), np.array([m]), rng)
        qs.append(q); rs.append(r)
    b = W.from_lists(qs, rs, (1, 1, 1, 1))
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, flag=0) as d:
        d.run(); d.run(); ms = d.stage_ms()
    print("n %7d pairs %5d cells %.2e fwd %.1f ms deciding %.1f  fwd GCUPS %.0f" % (n, pairs, b.cells, ms[0], ms[1], b.cells / ms[0] / 1e6))
