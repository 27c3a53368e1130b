"""GPU parity tests of the batched edit distance against the CPU checker
oracle/edit_oracle.c, through the C ABI (ssw_cuda_edit_distance_batch).  Integer results: bit-exact."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dist():
    from ciri_long_b200 import distance, ssw_wrap
    assert ssw_wrap.Aligner.libssw.ssw_cuda_device_count() > 0, "no CUDA device: the product has no CPU path"
    return distance


def oracle_batch(seqs, x_off, x_len, y_off, y_len):
    lib = C.CDLL(O.ORACLE_SO)
    out = np.zeros(len(x_len), dtype=np.int32)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    a = [np.ascontiguousarray(v, dtype=t) for v, t in ((x_off, np.int64), (x_len, np.int32), (y_off, np.int64), (y_len, np.int32))]
    lib.orc_edit_distance_batch(C.c_int32(len(out)), C.c_void_p(seqs.ctypes.data), C.c_void_p(a[0].ctypes.data),
                                C.c_void_p(a[1].ctypes.data), C.c_void_p(a[2].ctypes.data), C.c_void_p(a[3].ctypes.data),
                                C.c_void_p(out.ctypes.data))
    return out


def mutate(rng, s, rate):
    out = []
    for ch in s:
        u = rng.random()
        if u < rate / 3:
            continue
        out.append("ACGT"[rng.integers(0, 4)] if u < 2 * rate / 3 else ch)
        if rng.random() < rate / 3:
            out.append("ACGT"[rng.integers(0, 4)])
    return "".join(out)


def rand_seq(rng, n):
    return "".join(np.array(list("ACGT"))[rng.integers(0, 4, n)])


def test_known_answers(dist):
    pairs = [("kitten", "sitting", 3), ("flaw", "lawn", 2), ("intention", "execution", 5), ("GATTACA", "GCATGCU", 4),
             ("", "", 0), ("", "ACGT", 4), ("ACGT", "", 4), ("A", "A", 0), ("A", "C", 1), ("acgt", "ACGT", 4),
             ("ACGTN", "ACGTN", 0), ("NNNN", "ACGT", 4)]
    for x, y, d in pairs[:3]:                                   # English words: one call each (<= 16 symbols per batch)
        assert dist.distance(x, y) == d
    got = dist.distance_batch([p[0] for p in pairs[3:]], [p[1] for p in pairs[3:]])
    assert got.tolist() == [p[2] for p in pairs[3:]]


def test_every_kernel_instance_and_word_boundary(dist):
    """pattern lengths around every word / instance / tile boundary, related and unrelated texts"""
    rng = np.random.default_rng(5)
    xs, ys = [], []
    for m in (1, 2, 31, 32, 33, 50, 51, 63, 64, 65, 96, 127, 128, 129, 255, 256, 257, 300, 511, 512, 513, 700, 1023, 1024,
              1025, 1500, 2047, 2048, 2049, 3100):
        x = rand_seq(rng, m)
        for rate in (0.0, 0.1, 0.4):
            xs.append(x); ys.append(mutate(rng, x, rate))
        xs.append(x); ys.append(rand_seq(rng, int(m * 1.7) + 3))          # unrelated, longer
        xs.append(rand_seq(rng, max(1, m // 3))); ys.append(x)            # pattern is the second argument
        xs.append(x); ys.append(x[::-1])
    got = dist.distance_batch(xs, ys)
    exp = [O.edit_distance(x, y) for x, y in zip(xs, ys)]
    bad = [(len(x), len(y), int(g), e) for x, y, g, e in zip(xs, ys, got, exp) if g != e]
    assert not bad, bad[:5]


def test_junction_sized_batch(dist):
    """the avg_score shape (collapse.py:156-158): 20-nt genomic junction vs a 10-30 nt piece of the consensus"""
    rng = np.random.default_rng(6)
    n = 200000
    x_len = np.full(n, 20, dtype=np.int32)
    y_len = rng.integers(0, 31, n).astype(np.int32)
    total = int(x_len.sum() + y_len.sum())
    seqs = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.choice(5, total, p=[.24, .24, .24, .24, .04])]
    lens = np.stack([x_len, y_len], axis=1).reshape(-1).astype(np.int64)
    offs = np.cumsum(lens) - lens
    x_off, y_off = offs[0::2], offs[1::2]
    # make half of the pairs related: copy the start of x into y
    for p in range(0, n, 2):
        k = min(int(y_len[p]), 20)
        seqs[y_off[p]:y_off[p] + k] = seqs[x_off[p]:x_off[p] + k]
    got = dist.distance_arrays(seqs, x_off, x_len, y_off, y_len)
    exp = oracle_batch(seqs, x_off, x_len, y_off, y_len)
    assert np.array_equal(got, exp), np.flatnonzero(got != exp)[:10]


def test_cluster_distance_matrix(dist):
    """the O(k^2) loop of collapse.cluster_sequence (collapse.py:466-473) on read-like sequences"""
    rng = np.random.default_rng(7)
    base = rand_seq(rng, 900)
    seqs = [mutate(rng, base, 0.12) for _ in range(9)] + [rand_seq(rng, 400), base[:70], base[:40]]
    got = dist.cluster_distance_matrix(seqs)
    k = len(seqs)
    exp = np.zeros((k, k))
    for i in range(k):
        for j in range(i, k):
            exp[i][j] = O.edit_distance(seqs[i], seqs[j]) / max(len(seqs[i]), len(seqs[j]))
    exp = exp + exp.T
    assert np.array_equal(got, exp)


def test_properties_at_size(dist):
    """size-independent properties on long strings (no oracle): identity, symmetry, length bounds, and the
    exact distance of a string to itself with k substitutions planted far apart"""
    rng = np.random.default_rng(8)
    xs, ys, exact = [], [], []
    for n in (5000, 20000, 60000):
        x = rand_seq(rng, n)
        y = list(x)
        pos = rng.choice(n, 25, replace=False)
        for p in pos:
            y[p] = "ACGT"[("ACGT".index(y[p]) + 1) % 4]
        y = "".join(y)
        xs += [x, x, y, x]; ys += [x, y, x, x[: n // 2]]
        exact += [0, None, None, n - n // 2]
    got = dist.distance_batch(xs, ys).tolist()
    for k in range(0, len(xs), 4):
        assert got[k] == 0
        assert got[k + 1] == got[k + 2] and 0 < got[k + 1] <= 25
        assert got[k + 3] == exact[k + 3]


def test_too_many_symbols_is_refused(dist):
    from ciri_long_b200.ssw_wrap import SSWCudaError
    with pytest.raises(SSWCudaError):
        dist.distance_batch(["ABCDEFGHIJKLMNOPQRST"], ["abcdefghij"])


def test_curate_junction_batch(dist):
    """collapse.curate_junction (collapse.py:161-173) in two device batches against the reference's loop
    driven by the CPU checkers"""
    from ciri_long_b200 import callsites as cs
    oracle = O.Oracle()
    rng = np.random.default_rng(9)
    genome = rand_seq(rng, 400)
    st, en = 120, 300
    junc = mutate(rng, genome[en - 25:en] + genome[st:st + 25], 0.06)
    cands = []
    for i in range(st - 6, st + 6):
        for j in range(en - 6, en + 6):
            cands.append((i, j, genome[j - 10:j] + genome[i:i + 10]))
    got = cs.curate_junction_batch(cands, junc)
    exp = []
    mat = O.make_mat(10, 4)
    for i, j, tmp in cands:
        r = oracle.align(O.encode(junc), O.encode(tmp), mat, 8, 2, flag=1)
        x = junc[r["read_begin"]:r["read_end"]]
        exp.append((i, j, O.edit_distance(tmp, x) / len(tmp)))
    exp = sorted(exp, key=lambda t: t[2])
    assert got == exp
