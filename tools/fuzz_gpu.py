"""Randomised parity sweep: the CUDA path against the UNMODIFIED reference library (oracle/_ref/libssw.so)
on mixed shapes and random supported scoring schemes.  Test infrastructure (uses oracle/); prints one
line per scheme and the first mismatches.  argv: pairs per scheme, number of schemes, seed, family (mixed | edge | long | both | big), "flags" to vary flag and mask length."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multiprocessing as mp
import numpy as np
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
from oracle import oracle as O

_ref = None
def _init():
    global _ref
    _ref = O.RefLib()

def _one(job):
    q, r, mat, go, ge, flag, mask = job
    e = _ref.align(q, r, mat, go, ge, flag=flag, mask_len=mask)
    if e is None:                                   # the reference gave up (its traceback left the band, ssw.c:716-733)
        return None
    return (tuple(e[k] for k in O.FIELDS), tuple(int(c) for c in e["cigar"]))

def make_edge_pairs(n, rng):
    """degenerate inputs: one to three bases, homopolymers, all-N, identical sequences, query longer than reference"""
    qs, rs = [], []
    for k in range(n):
        kind = rng.integers(0, 7)
        if kind == 0:   q = rng.integers(0, 5, int(rng.integers(1, 4))); r = rng.integers(0, 5, int(rng.integers(1, 4)))
        elif kind == 1: b = int(rng.integers(0, 4)); q = np.full(int(rng.integers(1, 300)), b); r = np.full(int(rng.integers(1, 600)), b)
        elif kind == 2: q = np.full(int(rng.integers(1, 100)), 4); r = rng.integers(0, 5, int(rng.integers(1, 200)))
        elif kind == 3: q = rng.integers(0, 4, int(rng.integers(1, 1500))); r = q.copy()
        elif kind == 4: q = rng.integers(0, 4, int(rng.integers(200, 2000))); r = q[int(rng.integers(0, 100)):][:int(rng.integers(5, 60))].copy()
        elif kind == 5: r = rng.integers(0, 4, int(rng.integers(16, 40))); q = np.concatenate([r] * int(rng.integers(2, 30)))
        else:           q = rng.integers(0, 2, int(rng.integers(1, 500))); r = rng.integers(0, 2, int(rng.integers(1, 900)))      # two-letter alphabet: many ties
        if len(r) == 0: r = np.array([0])
        qs.append(q.astype(np.int8)); rs.append(r.astype(np.int8))
    return qs, rs


def make_long_pairs(n, rng):
    """references beyond 32 k columns (column-chunk tasks), queries from 20 bases to more than one tile"""
    qs, rs = [], []
    genome = rng.integers(0, 4, 400000).astype(np.int8)
    for k in range(n):
        nn = int(rng.integers(33000, 120000)); a = int(rng.integers(0, len(genome) - nn))
        r = genome[a:a + nn].copy()
        m = int(rng.choice([20, 50, 120, 250, 340, 400, 500, 800, 1100, 1500]))
        st = int(rng.integers(0, nn - m))
        rate = (0.02, 0.06, 0.12)[int(rng.integers(0, 3))]
        q, _ = W.noisy_channel(r[st:st + m].copy(), np.array([m]), rng, sub=rate, ins=rate, dele=rate, n_frac=0.01)
        if rng.random() < 0.3:      # a second copy of the query elsewhere: second-best score over merged column records
            st2 = int(rng.integers(0, nn - len(q))); r[st2:st2 + len(q)] = q
        qs.append(q); rs.append(r)
    return qs, rs


def make_pairs(n, rng, family="mixed"):
    if family == "edge": return make_edge_pairs(n, rng)
    if family == "long": return make_long_pairs(n, rng)
    if family == "big":                                     # scores around and beyond the 16-bit range (32-bit kernels, saturation)
        qs, rs = [], []
        for k in range(n):
            nn = int(rng.integers(2500, 4800)); r = rng.integers(0, 4, nn).astype(np.int8)
            m = int(rng.integers(2500, nn + 1)); st = int(rng.integers(0, nn - m + 1))
            rate = (0.0, 0.005, 0.02, 0.05)[int(rng.integers(0, 4))]
            q, _ = W.noisy_channel(r[st:st + m].copy(), np.array([m]), rng, sub=rate, ins=rate, dele=rate)
            qs.append(q); rs.append(r)
        return qs, rs
    if family == "both":                                    # short and long references in one batch, shuffled
        qa, ra = make_pairs(n - n // 10, rng, "mixed"); qb, rb = make_long_pairs(n // 10, rng)
        order = rng.permutation(n)
        qs, rs = qa + qb, ra + rb
        return [qs[i] for i in order], [rs[i] for i in order]
    qs, rs = [], []
    for k in range(n):
        shape = rng.integers(0, 6)
        if shape == 0:   m, nn = rng.integers(1, 80), rng.integers(1, 80)
        elif shape == 1: m, nn = rng.integers(100, 900), rng.integers(300, 3000)
        elif shape == 2: nn = rng.integers(50, 1200); m = max(1, nn + rng.integers(-30, 31))
        elif shape == 3: m, nn = rng.integers(200, 2500), rng.integers(15, 80)
        elif shape == 4: m, nn = rng.integers(900, 1300), rng.integers(1000, 1500)
        else:            m, nn = rng.integers(20, 400), rng.integers(3000, 9000)
        m, nn = int(m), int(nn)
        r = rng.integers(0, 4, nn).astype(np.int8)
        mode = rng.integers(0, 4)
        if mode == 0 or nn < 4:
            q = rng.integers(0, 4, m).astype(np.int8)
        else:
            L = min(m, nn); st = int(rng.integers(0, nn - L + 1))
            rate = (0.02, 0.06, 0.15)[int(rng.integers(0, 3))]
            q, _ = W.noisy_channel(r[st:st + L].copy(), np.array([L]), rng, sub=rate, ins=rate, dele=rate, max_run=int(rng.integers(1, 9)),
                                   n_frac=(0.0, 0.02)[int(rng.integers(0, 2))])
            if len(q) == 0: q = rng.integers(0, 4, 3).astype(np.int8)
            if mode == 2:   # periodic reference: many equal-scoring placements (tie-breaking)
                unit = r[:max(2, nn // int(rng.integers(2, 9)))]
                r = np.tile(unit, nn // len(unit) + 1)[:nn].copy()
        qs.append(q); rs.append(r)
    return qs, rs

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    schemes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    family = sys.argv[4] if len(sys.argv) > 4 else "mixed"
    vary = len(sys.argv) > 5 and sys.argv[5] == "flags"
    fixed = [(1, 1, 1, 1), (10, 4, 8, 2), (2, 2, 3, 1), (3, 2, 2, 2), (2, 1, 1, 1), (5, 4, 6, 2)]
    pool = mp.Pool(os.cpu_count(), initializer=_init)
    total_bad = 0
    for s in range(schemes):
        if s < len(fixed): p = fixed[s]
        else:
            ge = int(rng.integers(1, 6)); go = ge + int(rng.integers(0, 8)); mis = int(rng.integers(1, 2 * ge + 1)); mat = int(rng.integers(1, 12))
            p = (mat, mis, go, ge)
        qs, rs = make_pairs(n, rng, family)
        b = W.from_lists(qs, rs, p)
        # the ssw_align flag (ssw.c:779-869) and the mask length of the second-best search, when asked for
        flag = int(rng.choice([0, 1, 2, 4])) if vary else 1
        masks = None
        if vary:
            masks = np.array([O.default_mask_len(len(q)) if u < 0.4 else (15 if u < 0.6 else (int(rng.integers(1, 15)) if u < 0.8 else int(rng.integers(16, 2000))))
                              for q, u in zip(qs, rng.random(n))], dtype=np.int32)
        t0 = time.perf_counter()
        with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, p[0], p[1], p[2], p[3], flag=flag, mask_len=masks) as d:
            d.run(); rec, cig = d.fetch()
        t1 = time.perf_counter()
        mat = O.make_mat(p[0], p[1])
        exp = pool.map(_one, [(qs[i], rs[i], mat, p[2], p[3], flag, None if masks is None else int(masks[i])) for i in range(n)], chunksize=64)
        t2 = time.perf_counter()
        bad = []; ub = 0
        for i in range(n):
            r = rec[i]
            if exp[i] is None:
                if (r["status"] & 0xff) == 0: bad.append((i, "reference returned NULL, device did not", len(qs[i]), len(rs[i])))
                continue
            got = (int(r["score1"]), int(r["score2"]), int(r["ref_begin1"]), int(r["ref_end1"]), int(r["read_begin1"]), int(r["read_end1"]), int(r["ref_end2"]))
            gc = tuple(int(c) for c in cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]])
            if (r["status"] & 0xff) == 1 and got == exp[i][0]:
                # the traceback leaves the band: the reference goes on reading direction bytes it never wrote
                # (ssw.c:643-644 with j < beg) and returns whatever that memory yields; the device reports the pair
                ub += 1
                continue
            ok = (r["status"] & 0xff) == 0 and got == exp[i][0] and gc == exp[i][1]
            if not ok and not (exp[i][0][0] == 0):          # score 0: the reference reads ref[-1] (undefined), fields still compared above
                bad.append((i, int(r["status"]), len(qs[i]), len(rs[i]), got, exp[i][0], gc == exp[i][1]))
            elif not ok and got != exp[i][0]:
                bad.append((i, int(r["status"]), len(qs[i]), len(rs[i]), got, exp[i][0], gc == exp[i][1]))
        total_bad += len(bad)
        print("scheme %s flag %d pairs %d cells %.2e gpu %.2fs ref(%d cores) %.1fs mismatches %d traceback-left-band %d %s" % (p, flag, n, b.cells, t1 - t0, os.cpu_count(), t2 - t1, len(bad), ub, bad[:3]), flush=True)
    print("TOTAL mismatches", total_bad)

if __name__ == "__main__":
    main()
