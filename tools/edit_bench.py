"""Measure the batched edit distance through the C ABI, host buffers in and out.

Two shapes: the avg_score pairs of collapse.curate_junction (20-nt junction vs a <= 30-nt piece, millions of
pairs) and the all-pairs loop of collapse.cluster_sequence (homopolymer-compressed reads of ~1 kb).
The CPU figure beside it is the checker oracle/edit_oracle.c (scalar O(m*n) port, one core) on a sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import ciri_long_b200
from ciri_long_b200 import distance as D
from oracle import oracle as O


def junction_pairs(n, rng):
    x_len = np.full(n, 20, dtype=np.int32)
    y_len = rng.integers(10, 31, n).astype(np.int32)
    total = int(x_len.sum() + y_len.sum())
    seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, total)]
    lens = np.stack([x_len, y_len], axis=1).reshape(-1).astype(np.int64)
    offs = np.cumsum(lens) - lens
    return seqs, offs[0::2].copy(), x_len, offs[1::2].copy(), y_len


def cluster_pairs(k, length, rng):
    base = rng.integers(0, 4, length)
    seqs, off, ln = [], [], []
    pos = 0
    for _ in range(k):
        s = base.copy()
        hit = rng.random(length) < 0.1
        s[hit] = rng.integers(0, 4, int(hit.sum()))
        s = s[rng.random(length) > 0.05]
        seqs.append(s); off.append(pos); ln.append(len(s)); pos += len(s)
    seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[np.concatenate(seqs)]
    ii, jj = np.triu_indices(k)
    off, ln = np.array(off, dtype=np.int64), np.array(ln, dtype=np.int32)
    return seqs, off[ii], ln[ii], off[jj], ln[jj]


def oracle_time(args, sample):
    lib = C.CDLL(O.ORACLE_SO)
    seqs, xo, xl, yo, yl = args
    idx = np.arange(min(sample, len(xl)))
    a = [np.ascontiguousarray(v[idx]) for v in (xo, xl, yo, yl)]
    out = np.zeros(len(idx), dtype=np.int32)
    t0 = time.perf_counter()
    lib.orc_edit_distance_batch(C.c_int32(len(idx)), C.c_void_p(seqs.ctypes.data), C.c_void_p(a[0].ctypes.data),
                                C.c_void_p(a[1].ctypes.data), C.c_void_p(a[2].ctypes.data), C.c_void_p(a[3].ctypes.data),
                                C.c_void_p(out.ctypes.data))
    dt = time.perf_counter() - t0
    cells = float((a[1].astype(np.int64) * a[3].astype(np.int64)).sum())
    return out, cells / dt / 1e9, len(idx), dt


def run(name, args, sample, reps=3):
    seqs, xo, xl, yo, yl = args
    cells = float((xl.astype(np.int64) * yl.astype(np.int64)).sum())
    D.distance_arrays(seqs, xo, xl, yo, yl)                              # warm-up (context, pools)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        got = D.distance_arrays(seqs, xo, xl, yo, yl)
        best = min(best, time.perf_counter() - t0)
    exp, cpu_gcups, ns, dt = oracle_time(args, sample)
    assert np.array_equal(got[:ns], exp)
    return dict(shape=name, pairs=int(len(xl)), cells=cells, e2e_ms=best * 1e3, e2e_pairs_per_s=len(xl) / best,
                e2e_gcups=cells / best / 1e9,
                cpu_checker=dict(kind="port", cores=1, gcups=cpu_gcups, sample="%d pairs, %.1f s" % (ns, dt)),
                parity="%d sampled pairs identical to the checker" % ns)


if __name__ == "__main__":
    rng = np.random.default_rng(20261017)
    res = [run("junction 20 nt vs 10-30 nt (collapse.py:156-158)", junction_pairs(4_000_000, rng), 400_000),
           run("cluster all-pairs, 300 reads of ~1 kb (collapse.py:466-473)", cluster_pairs(300, 1000, rng), 3000),
           run("cluster all-pairs, 120 reads of ~3 kb", cluster_pairs(120, 3000, rng), 600)]
    print(json.dumps(dict(metric="batched edit distance through ssw_cuda_edit_distance_batch (host buffers)", results=res)))
