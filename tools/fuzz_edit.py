"""Randomised parity sweep of the batched edit distance against the CPU checker (oracle/edit_oracle.c).
Test infrastructure.  argv: pairs, seed."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ciri_long_b200
from ciri_long_b200 import distance as D
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
alphabet = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
chunks, xo, xl, yo, yl = [], [], [], [], []
pos = 0
for k in range(n):
    kind = rng.integers(0, 5)
    if kind == 0:   m = int(rng.integers(0, 70))
    elif kind == 1: m = int(rng.integers(60, 140))
    elif kind == 2: m = int(rng.integers(120, 600))
    elif kind == 3: m = int(rng.integers(500, 1100))
    else:           m = int(rng.integers(1000, 2600))
    x = alphabet[rng.integers(0, 4 if rng.random() < 0.8 else 9, m)]
    mode = rng.integers(0, 3)
    if mode == 0:
        y = alphabet[rng.integers(0, 4, int(rng.integers(0, max(2, 2 * m))))]
    else:
        rate = (0.02, 0.1, 0.3)[int(rng.integers(0, 3))]
        keep = rng.random(m) > rate
        y = x[keep].copy()
        sub = rng.random(len(y)) < rate
        y[sub] = alphabet[rng.integers(0, 4, int(sub.sum()))]
        if rng.random() < 0.5 and len(y):
            ins = rng.integers(0, len(y), int(rate * len(y)))
            y = np.insert(y, ins, alphabet[rng.integers(0, 4, len(ins))])
        if mode == 2 and len(y) > 4:
            y = y[int(rng.integers(0, len(y) // 2)):]
    if rng.random() < 0.5: x, y = y, x
    chunks += [x, y]
    xo.append(pos); xl.append(len(x)); pos += len(x)
    yo.append(pos); yl.append(len(y)); pos += len(y)
seqs = np.concatenate(chunks + [np.zeros(1, dtype=np.uint8)])
xo, yo = np.array(xo, dtype=np.int64), np.array(yo, dtype=np.int64)
xl, yl = np.array(xl, dtype=np.int32), np.array(yl, dtype=np.int32)
t0 = time.perf_counter(); got = D.distance_arrays(seqs, xo, xl, yo, yl); t1 = time.perf_counter()
lib = C.CDLL(O.ORACLE_SO)
exp = np.zeros(n, dtype=np.int32)
lib.orc_edit_distance_batch(C.c_int32(n), C.c_void_p(seqs.ctypes.data), C.c_void_p(xo.ctypes.data), C.c_void_p(xl.ctypes.data),
                            C.c_void_p(yo.ctypes.data), C.c_void_p(yl.ctypes.data), C.c_void_p(exp.ctypes.data))
t2 = time.perf_counter()
bad = np.flatnonzero(got != exp)
print("pairs %d cells %.2e gpu %.2fs checker %.1fs mismatches %d %s" % (n, float((xl.astype(np.int64) * yl).sum()), t1 - t0, t2 - t1, len(bad),
      [(int(i), int(xl[i]), int(yl[i]), int(got[i]), int(exp[i])) for i in bad[:5]]))
