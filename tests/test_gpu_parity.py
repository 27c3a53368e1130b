"""GPU parity tests: the CUDA path, called through the C ABI of libssw_cuda.so, against the CPU oracle
(oracle/ssw_oracle.c) and the golden vectors generated from the unmodified reference.  Bit-exact on
every field: score, ref/read begin/end, score2, ref_end2, CIGAR ops."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sw():
    from ciri_long_b200 import ssw_wrap
    if ssw_wrap.Aligner.libssw.ssw_cuda_device_count() <= 0:
        pytest.skip("no CUDA device: the product has no CPU path")
    return ssw_wrap


def run_batch(sw, b, flag=1):
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, b.match, b.mismatch, b.gap_open,
                        b.gap_extend, flag=flag) as d:
        d.run()
        return d.fetch()


def check_batch(sw, oracle, b, flag=1, sample=None):
    rec, cig = run_batch(sw, b, flag)
    mat = O.make_mat(b.match, b.mismatch)
    idx = range(len(b)) if sample is None else sample
    bad = []
    for i in idx:
        exp = oracle.align(b.query(i), b.ref(i), mat, b.gap_open, b.gap_extend, flag=flag)
        r = rec[i]
        got = dict(score=int(r["score1"]), score2=int(r["score2"]), ref_begin=int(r["ref_begin1"]),
                   ref_end=int(r["ref_end1"]), read_begin=int(r["read_begin1"]), read_end=int(r["read_end1"]),
                   ref_end2=int(r["ref_end2"]),
                   cigar=cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]].tolist())
        if (r["status"] & 0xff) != 0 or not O.same(got, exp) or int(r["word"]) != exp["word"]:
            bad.append((i, int(r["status"]), int(r["word"]), {k: got[k] for k in O.FIELDS},
                        {k: exp[k] for k in O.FIELDS}, exp["word"], got["cigar"] == exp["cigar"]))
    assert not bad, "%s: %d mismatches, first: %s" % (b.name, len(bad), bad[:3])
    return rec, cig


def test_golden_cases(sw, golden):
    """every committed golden vector (fuzz, overflow boundary, planted repeats, G5-G7) per parameter set"""
    from ciri_long_b200 import workloads as W
    by_params = {}
    for c in golden["cases"]:
        by_params.setdefault(tuple(c["params"]), []).append(c)
    for p, cases in by_params.items():
        b = W.from_lists([O.encode(c["query"]) for c in cases], [O.encode(c["ref"]) for c in cases], p)
        rec, cig = run_batch(sw, b)
        for i, c in enumerate(cases):
            e, r = c["expected"], rec[i]
            got = (int(r["score1"]), int(r["score2"]), int(r["ref_begin1"]), int(r["ref_end1"]),
                   int(r["read_begin1"]), int(r["read_end1"]), int(r["ref_end2"]))
            want = tuple(e[k] for k in O.FIELDS)
            assert (r["status"] & 0xff) == 0, (c["name"], int(r["status"]))
            assert got == want, (c["name"], got, want)
            assert cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]].tolist() == e["cigar"], c["name"]


def test_golden_testfa(sw, golden):
    """tests/test.fa of the reference: 437 nt vs 430,314 nt, both orientations, both parameter sets"""
    from ciri_long_b200 import workloads as W
    for c in golden["testfa"]:
        b = W.from_lists([golden["seqs"][c["query"]]], [golden["seqs"][c["ref"]]], tuple(c["params"]))
        rec, cig = run_batch(sw, b)
        e, r = c["expected"], rec[0]
        got = (int(r["score1"]), int(r["score2"]), int(r["ref_begin1"]), int(r["ref_end1"]),
               int(r["read_begin1"]), int(r["read_end1"]), int(r["ref_end2"]))
        assert (r["status"] & 0xff) == 0, (c["name"], int(r["status"]))
        assert got == tuple(e[k] for k in O.FIELDS), (c["name"], got)
        assert cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]].tolist() == e["cigar"], c["name"]


@pytest.mark.parametrize("params", [(1, 1, 1, 1), (10, 4, 8, 2), (2, 2, 3, 1), (2, 2, 2, 2)])
def test_bsj_refinement_shape(sw, oracle, params):
    from ciri_long_b200 import workloads as W
    check_batch(sw, oracle, W.bsj_refinement_pairs(96, seed=11, params=params))


def test_rolling_circle_shape(sw, oracle):
    from ciri_long_b200 import workloads as W
    check_batch(sw, oracle, W.rolling_circle_pairs(24, seed=12, read_min=800, read_max=3000))


@pytest.mark.parametrize("length", [64, 128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("params", [(1, 1, 1, 1), (10, 4, 8, 2)])
def test_square_sweep(sw, oracle, length, params):
    from ciri_long_b200 import workloads as W
    n = max(4, 2048 // length)
    check_batch(sw, oracle, W.square_pairs(n, length, params=params))
    check_batch(sw, oracle, W.square_pairs(n, length, params=params), flag=0)


def test_tiny_junction_pairs(sw, oracle):
    from ciri_long_b200 import workloads as W
    check_batch(sw, oracle, W.junction_pairs(2000, seed=13))


def test_overflow_boundary(sw, oracle):
    from ciri_long_b200 import workloads as W
    check_batch(sw, oracle, W.overflow_boundary_pairs())
    check_batch(sw, oracle, W.overflow_boundary_pairs(params=(2, 2, 2, 2), lengths=range(118, 134)))
    check_batch(sw, oracle, W.overflow_boundary_pairs(params=(10, 4, 8, 2), lengths=range(20, 30)))


def test_ragged_and_degenerate(sw, oracle):
    """1-base sequences, all-N, no match at all, query longer than reference"""
    from ciri_long_b200 import workloads as W
    qs = [np.array([0], np.int8), np.array([4, 4, 4, 4], np.int8), np.zeros(24, np.int8) + 1,
          np.arange(200, dtype=np.int8) % 4, np.array([2, 3], np.int8)]
    rs = [np.array([0], np.int8), np.array([0, 1, 2, 3], np.int8), np.zeros(20, np.int8),
          np.arange(37, dtype=np.int8) % 4, np.arange(3000, dtype=np.int8) % 4]
    for p in [(1, 1, 1, 1), (10, 4, 8, 2)]:
        check_batch(sw, oracle, W.from_lists(qs, rs, p, name="degenerate"))


def test_large_batch_properties(sw, oracle):
    """BASELINE-sized shape at reduced count: sampled oracle parity + size-independent invariants"""
    from ciri_long_b200 import workloads as W
    b = W.bsj_refinement_pairs(20000, seed=21)
    rec, cig = check_batch(sw, oracle, b, sample=range(0, 20000, 400))
    assert ((rec["status"] & 0xff) == 0).all()
    # CIGAR consistency: M+I covers the aligned query span, M+D the reference span (1/1/1/1: no dropped deletions)
    ops = cig & 0xF
    lens = (cig >> 4).astype(np.int64)
    seg = np.repeat(np.arange(len(b)), rec["cigar_len"])
    order = np.argsort(rec["cigar_off"], kind="stable")
    assert (np.diff(rec["cigar_off"][order]) == rec["cigar_len"][order][:-1]).all()      # dense, no overlap
    seg = np.repeat(order, rec["cigar_len"][order])
    qspan = np.bincount(seg, weights=np.where(ops != 2, lens, 0), minlength=len(b))
    rspan = np.bincount(seg, weights=np.where(ops != 1, lens, 0), minlength=len(b))
    assert (qspan == rec["read_end1"] - rec["read_begin1"] + 1).all()
    assert (rspan == rec["ref_end1"] - rec["ref_begin1"] + 1).all()
    # running the same resident batch twice gives identical results (idempotence of ssw_batch_run)
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1) as d:
        d.run(); r1, c1 = d.fetch()
        d.run(); r2, c2 = d.fetch()
    for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len"):
        assert (r1[k] == r2[k]).all() and (r1[k] == rec[k]).all()


def test_wrapper_drop_in(sw, golden):
    """Aligner.align / align_batch / align_pairs mirror the reference wrapper (incl. soft-clipped CIGAR string)"""
    cases = golden["cases"][:40]
    for c in cases[:8]:
        p = c["params"]
        al = sw.Aligner(c["ref"], match=p[0], mismatch=p[1], gap_open=p[2], gap_extend=p[3],
                        report_secondary=True, report_cigar=True)
        res = al.align(c["query"])
        e = c["expected"]
        assert (res.score, res.ref_begin, res.ref_end, res.query_begin, res.query_end) == \
            (e["score"], e["ref_begin"], e["ref_end"], e["read_begin"], e["read_end"])
        assert (res.score2 or 0) == e["score2"] and res.cigar_string == e["cigar_string"]
        assert al.align(c["query"], min_score=10 ** 6) is None
    p = tuple(cases[0]["params"])
    same = [c for c in cases if tuple(c["params"]) == p]
    out = sw.align_pairs([c["ref"] for c in same], [c["query"] for c in same], *p, report_secondary=True,
                         report_cigar=True)
    for c, res in zip(same, out):
        assert res.cigar_string == c["expected"]["cigar_string"] and res.score == c["expected"]["score"]
    al = sw.Aligner(same[0]["ref"], *p, report_cigar=False)
    out = al.align_batch([c["query"] for c in same])
    ref0 = [sw.Aligner(same[0]["ref"], *p).align(c["query"]) for c in same[:5]]
    for a, b_ in zip(out[:5], ref0):
        assert (a.score, a.ref_begin, a.query_end) == (b_.score, b_.ref_begin, b_.query_end) and a.cigar_string is None


def test_one_shot_chunked_call(sw, oracle, monkeypatch):
    """ssw_align_batch cuts large batches into chunks on alternating streams: same results as one resident batch"""
    from ciri_long_b200 import workloads as W
    b = W.bsj_refinement_pairs(700, seed=31)
    rec, cig = run_batch(sw, b)
    monkeypatch.setenv("SSW_CUDA_CHUNK", "128")
    r2, c2 = sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, b.match, b.mismatch, b.gap_open, b.gap_extend)
    for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len", "status"):
        assert (rec[k] == r2[k]).all(), k
    for i in range(len(b)):
        a = cig[rec["cigar_off"][i]:rec["cigar_off"][i] + rec["cigar_len"][i]]
        c = c2[r2["cigar_off"][i]:r2["cigar_off"][i] + r2["cigar_len"][i]]
        assert (a == c).all()
    assert len(c2) == len(cig)
    # a too small caller buffer is reported, not overrun
    with pytest.raises(sw.SSWCudaError):
        sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, cig=np.empty(10, np.uint32))


def test_scores_beyond_16_bit_comfort(sw, oracle):
    """Pairs whose score reaches the reference's int16 saturation (ssw.c:442) or the packed kernel's range
    limit are re-done by the 32-bit kernels: still bit-exact, including the saturated results."""
    from ciri_long_b200 import workloads as W
    # match = 10: 3.4 kb and 3.6 kb perfect matches saturate at 32767; 4096-nt noisy pairs score ~33k
    b = W.square_pairs(2, 3400, params=(10, 4, 8, 2), noise=False)
    check_batch(sw, oracle, b)
    b = W.square_pairs(1, 3600, params=(10, 4, 8, 2), noise=False)
    check_batch(sw, oracle, b)
    check_batch(sw, oracle, W.square_pairs(2, 4096, params=(10, 4, 8, 2)))
    # gap_open == gap_extend: scores above 16000 leave the packed truncated-F kernel
    check_batch(sw, oracle, W.square_pairs(1, 8400, params=(2, 2, 2, 2), noise=False))


def test_batched_call_sites(sw, oracle):
    """callsites.py: the two-phase versions of find_bsj.align_clip_segments / collapse.junc_score give what
    a per-pair loop through the drop-in Aligner gives"""
    from ciri_long_b200 import callsites as cs
    rng = np.random.default_rng(77)
    bases = np.array(list("ACGT"))
    genome = "".join(bases[rng.integers(0, 4, 6000)])
    items = []
    for k in range(24):
        a = int(rng.integers(500, 4000)); L = int(rng.integers(300, 700))
        circ_src = genome[a:a + L]
        cut = int(rng.integers(30, 120))
        circ = circ_src[cut:] + circ_src[:cut]                    # rotated: the tail of the read maps before its head
        strand = 1 if k % 3 else -1
        if strand < 0:
            circ = cs.revcomp(circ)
        q_st, q_en = (0, L - cut) if k % 2 else (5, L - cut)
        hit = cs.Hit(q_st, q_en, a + cut, a + L, strand)
        tmp_start, tmp_end = max(hit.r_st - 2000, 0), min(hit.r_en + 2000, len(genome))
        items.append((circ, hit, genome[tmp_start:tmp_end], tmp_start, tmp_end))
    items.append(("ACGTACGTACGTACGTACGTAAAA", cs.Hit(2, 20, 100, 118, 1), None, 0, 0))     # < 20 clipped bases: no alignment
    got = cs.align_clip_segments_batch(items)
    for it, g in zip(items, got):
        assert g == cs.align_clip_segments_batch([it])[0]
    # one item by hand, through the per-call drop-in
    circ, hit, window, tmp_start, tmp_end = items[1]
    clip_seq = circ[hit.q_en:] + circ[:hit.q_st]
    res = sw.Aligner(window if hit.strand > 0 else cs.revcomp(window), 1, 1, 1, 1).align(clip_seq)
    if hit.strand > 0:
        exp = (tmp_start + res.ref_begin, tmp_start + res.ref_end)
    else:
        exp = (tmp_end - res.ref_end, tmp_end - res.ref_begin)
    assert got[1][3][:2] == exp
    assert got[-1] == ("GTACGTACGTACGTACGTAAAAAC", 99, 118, (None, None, 6))
    spans = [genome[100:400], genome[1000:1200]]
    reads = [[genome[380:400] + genome[100:130], genome[385:400] + genome[100:140]], [genome[1180:1200] + genome[1000:1030]]]
    sc = cs.junc_score_batch(spans, reads)
    for span, rs, s in zip(spans, reads, sc):
        al = sw.Aligner(span * 2, 10, 4, 8, 2)
        assert abs(s - np.mean([al.align(r).score for r in rs])) < 1e-9


class _OracleAlignment:
    """what the reference's PyAlignRes exposes, filled from the CPU oracle (independent of the CUDA path)"""
    def __init__(self, oracle, ref, query, p):
        r = oracle.align(O.encode(query), O.encode(ref), O.make_mat(p[0], p[1]), p[2], p[3], flag=1)
        self.score, self.ref_begin, self.ref_end = r["score"], r["ref_begin"], r["ref_end"]
        self.query_begin, self.query_end = r["read_begin"], r["read_end"]
        self.cigar_string = O.cigar_string(r, len(query))


def test_batched_collapse_call_sites(sw, oracle):
    """callsites.py: correct_cluster head / refined sequences / exon_score in two-phase form against the
    reference's per-read loops (collapse.py:251-265, 369-385, 760-774) driven by the CPU oracle"""
    from ciri_long_b200 import callsites as cs
    from ciri_long_b200.workloads import noisy_channel
    rng = np.random.default_rng(99)
    bases = np.array(list("ACGT"))
    P = (10, 4, 8, 2)

    def noisy(seq):
        codes, _ = noisy_channel(O.encode(seq), np.array([len(seq)]), rng)
        return "".join("ACGTN"[c] for c in codes)

    clusters = []
    for c in range(6):
        L = int(rng.integers(260, 900))
        circ = "".join(bases[rng.integers(0, 4, L)])
        reads = []
        for _ in range(int(rng.integers(2, 7))):
            rot = int(rng.integers(0, L))
            reads.append(noisy(circ[rot:] + circ[:rot]))
        clusters.append((reads[0], reads[1:]))
    got = cs.cluster_junction_seqs_batch(clusters)
    for (ref_seq, qs), (template, juncs) in zip(clusters, got):
        head_pos = [_OracleAlignment(oracle, ref_seq[:50], q, P).ref_begin for q in qs]
        exp_template = cs.transform_seq(ref_seq, max(head_pos))
        exp = [cs.get_junc_seq(exp_template, -max(head_pos) // 2, 25)]
        for q in qs:
            al = _OracleAlignment(oracle, exp_template, q, P)
            exp.append(cs.get_junc_seq(cs.transform_seq(q, al.query_begin), -max(head_pos) // 2, 25))
        assert template == exp_template and juncs == exp

    items = []
    for c in range(5):
        L = int(rng.integers(200, 700))
        circ = "".join(bases[rng.integers(0, 4, L)])
        junc = circ[-25:] + circ[:25]                                 # genome_junction_seq: 25 nt either side of the BSJ
        reads = []
        for k in range(int(rng.integers(1, 6))):
            rot = int(rng.integers(0, L))
            reads.append(("read%d_%d" % (c, k), noisy(circ[rot:] + circ[:rot])))
        reads.append(("unrelated%d" % c, "".join(bases[rng.integers(0, 4, 120)])))
        items.append((junc, reads))
    got = cs.refined_sequences_batch(items)
    n_rotated = 0
    for (junc, reads), res in zip(items, got):
        for (rid, seq), g in zip(reads, res):
            al = _OracleAlignment(oracle, junc, seq * 2, P)
            pos = cs.find_alignment_pos(al, len(junc) // 2)
            exp = (rid, seq) if pos is None else (rid, cs.transform_seq(seq, pos % len(seq)))
            n_rotated += pos is not None
            assert g == exp
    assert n_rotated > 5

    cons = "".join(bases[rng.integers(0, 4, 900)])
    exon_pairs = [noisy(cons[a:a + 120] + cons[b:b + 150]) for a, b in ((0, 300), (100, 500), (350, 700), (20, 200))]
    exon_pairs.append(cs.revcomp(cons[50:400]))
    got = cs.exon_scores_batch(cons, exon_pairs)
    exp = []
    for q in exon_pairs:
        al = _OracleAlignment(oracle, cons, q, P)
        exp.append(al.ref_end - al.ref_begin)
    assert got == exp
    assert cs.exon_scores_batch(cons, []) == []


def test_long_references_in_column_chunks(sw, oracle):
    """find_bsj window shape (find_bsj.py:182-233): short queries against references of tens to hundreds of
    kilobases.  The forward pass runs as column-chunk tasks (ChunkPlan), the reverse pass as a bounded first
    look plus chunk tasks for the pairs without a stop column; every field incl. the second-best score
    (which reads the merged column records) and the CIGAR must come out as for whole pairs"""
    from ciri_long_b200 import workloads as W
    rng = np.random.default_rng(123)
    # (3,2,2,2) / (2,1,1,1): gap_open == gap_extend with match != mismatch, where the reverse pass can jump over
    # score1 without meeting it and then scans (here: in column-chunk tasks) the whole prefix
    for params in ((1, 1, 1, 1), (10, 4, 8, 2), (3, 2, 2, 2), (2, 1, 1, 1)):
        qs, rs = [], []
        for k in range(14):
            n = int(rng.integers(33000, 140000))
            r = rng.integers(0, 4, n).astype(np.int8)
            m = int((20, 60, 150, 300, 340, 420, 500, 700, 1000, 1100, 1500, 250, 480, 90)[k])
            st = int(rng.integers(0, n - m))
            q, _ = W.noisy_channel(r[st:st + m].copy(), np.array([m]), rng, n_frac=0.01)
            if k % 5 == 0:                                            # a second, weaker copy far away: score2 / ref_end2
                st2 = (st + n // 2) % (n - m)
                q2, _ = W.noisy_channel(q.copy(), np.array([len(q)]), rng, sub=0.15)
                r = r.copy(); r[st2:st2 + min(len(q2), n - st2)] = q2[:min(len(q2), n - st2)]
            if k == 3:                                                # alignment across a chunk border
                st = 8192 * 4 - m // 2
                q, _ = W.noisy_channel(r[st:st + m].copy(), np.array([m]), rng)
            qs.append(q); rs.append(r)
        qs.append(rng.integers(0, 4, 200).astype(np.int8)); rs.append(rng.integers(0, 4, 70000).astype(np.int8))   # unrelated
        # last chunk shorter than the 64-strip pipeline (chunks of 8192 columns in a batch this small): its columns
        # must still reach the merged column records that the second-best scan reads
        for tail in (1, 15, 63, 64, 100):
            n = 8192 * 5 + tail
            r = rng.integers(0, 4, n).astype(np.int8)
            q, _ = W.noisy_channel(r[5000:5300].copy(), np.array([300]), rng)
            r[n - len(q) - 3:n - 3] = q                               # a perfect copy that ends in the last chunk
            qs.append(q); rs.append(r)
        b = W.from_lists(qs, rs, params, name="long-ref %s" % (params,))
        check_batch(sw, oracle, b)


def test_randomized_mixture(sw, oracle):
    """a small edition of tools/fuzz_gpu.py (which sweeps against the reference library itself): mixed shapes incl.
    periodic references (equal-score ties), queries of more than one strip tile and degenerate inputs, three
    scoring schemes per run, every field and CIGAR against the oracle"""
    from ciri_long_b200 import workloads as W
    rng = np.random.default_rng(2026)
    for params in ((1, 1, 1, 1), (2, 1, 1, 1), (5, 5, 10, 3)):
        qs, rs = [], []
        for k in range(260):
            shape = k % 7
            if shape == 0:   m, n = int(rng.integers(1, 60)), int(rng.integers(1, 60))
            elif shape == 1: m, n = int(rng.integers(100, 700)), int(rng.integers(300, 2500))
            elif shape == 2: n = int(rng.integers(50, 900)); m = max(1, n + int(rng.integers(-30, 31)))
            elif shape == 3: m, n = int(rng.integers(200, 1800)), int(rng.integers(15, 80))
            elif shape == 4: m, n = int(rng.integers(1030, 1300)), int(rng.integers(1000, 1400))
            elif shape == 5: m, n = int(rng.integers(20, 300)), int(rng.integers(2000, 6000))
            else:            m, n = int(rng.integers(1, 400)), int(rng.integers(1, 700))
            if shape == 6:
                r = rng.integers(0, 2, n).astype(np.int8); q = rng.integers(0, 2, m).astype(np.int8)       # two letters: ties everywhere
            else:
                r = rng.integers(0, 4, n).astype(np.int8)
                if k % 3 == 2:                                       # periodic reference
                    unit = r[:max(2, n // int(rng.integers(2, 9)))]
                    r = np.tile(unit, n // len(unit) + 1)[:n].copy()
                L = min(m, n); st = int(rng.integers(0, n - L + 1))
                rate = (0.02, 0.06, 0.15)[k % 3]
                q, _ = W.noisy_channel(r[st:st + L].copy(), np.array([L]), rng, sub=rate, ins=rate, dele=rate, n_frac=0.01 * (k % 2))
                if len(q) == 0:
                    q = rng.integers(0, 4, 3).astype(np.int8)
            qs.append(q); rs.append(r)
        b = W.from_lists(qs, rs, params, name="randomized %s" % (params,))
        rec, cig = run_batch(sw, b)
        mat = O.make_mat(params[0], params[1])
        bad = []
        for i in range(len(b)):
            exp = oracle.align(b.query(i), b.ref(i), mat, params[2], params[3], flag=1)
            r = rec[i]
            if exp is None:                                          # traceback left the band: both sides must say so
                if (r["status"] & 0xff) == 0:
                    bad.append((i, "oracle: traceback error, device: ok"))
                continue
            got = dict(score=int(r["score1"]), score2=int(r["score2"]), ref_begin=int(r["ref_begin1"]), ref_end=int(r["ref_end1"]),
                       read_begin=int(r["read_begin1"]), read_end=int(r["read_end1"]), ref_end2=int(r["ref_end2"]),
                       cigar=cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]].tolist())
            if (r["status"] & 0xff) != 0 or not O.same(got, exp):
                bad.append((i, int(r["status"]), len(qs[i]), len(rs[i]), {k: got[k] for k in O.FIELDS}, {k: exp[k] for k in O.FIELDS}))
        assert not bad, "%s: %d mismatches, first: %s" % (b.name, len(bad), bad[:3])


@pytest.mark.parametrize("device_encode", ["0", "1"])
def test_string_batches(sw, monkeypatch, device_encode):
    """align_pairs on str inputs: one join and one table pass on the host, or (SSW_CUDA_DEVICE_ENCODE=1) the raw
    letters uploaded and converted on the device (ssw_wrap.py:234-252: A C G T N in either case, anything else
    N); same results as on pre-encoded arrays"""
    monkeypatch.setenv("SSW_CUDA_DEVICE_ENCODE", device_encode)
    rng = np.random.default_rng(5)
    letters = np.array(list("ACGTacgtNnXR-*"))
    p = [0.2, 0.2, 0.2, 0.2, 0.03, 0.03, 0.03, 0.03, 0.02, 0.02, 0.01, 0.01, 0.01, 0.01]
    refs, qs = [], []
    for k in range(300):
        n = int(rng.integers(1, 1500)); r = "".join(letters[rng.choice(len(letters), n, p=p)])
        if k % 2:
            st = int(rng.integers(0, n)); q = r[st:st + int(rng.integers(1, 400))].swapcase()
        else:
            q = "".join(letters[rng.choice(len(letters), int(rng.integers(1, 300)), p=p)])
        refs.append(r); qs.append(q)
    a = sw.align_pairs(refs, qs, 2, 2, 3, 1, report_secondary=True, report_cigar=True)
    b = sw.align_pairs([sw.encode_dna(x) for x in refs], [sw.encode_dna(x) for x in qs], 2, 2, 3, 1, report_secondary=True, report_cigar=True)
    for x, y in zip(a, b):
        assert (x is None) == (y is None)
        if x is not None:
            assert (x.score, x.score2, x.ref_begin, x.ref_end, x.query_begin, x.query_end, x.ref_end2, x.cigar_string) == \
                   (y.score, y.score2, y.ref_begin, y.ref_end, y.query_begin, y.query_end, y.ref_end2, y.cigar_string)
    # the record form of the same call
    rec, cig = sw.align_pairs(refs, qs, 2, 2, 3, 1, report_cigar=True, as_records=True)
    for x, r in zip(a, rec):
        if x is not None:
            assert (x.score, x.ref_begin, x.ref_end, x.query_begin, x.query_end) == \
                   (int(r["score1"]), int(r["ref_begin1"]), int(r["ref_end1"]), int(r["read_begin1"]), int(r["read_end1"]))
    # one reference, many queries (Aligner.align_batch)
    al = sw.Aligner(refs[0], 2, 2, 3, 1, report_cigar=True)
    got = al.align_batch(qs[:40])
    for q, g in zip(qs[:40], got):
        e = al.align(q)
        assert (g.score, g.ref_begin, g.ref_end, g.query_begin, g.query_end, g.cigar_string) == \
               (e.score, e.ref_begin, e.ref_end, e.query_begin, e.query_end, e.cigar_string)


def test_mixed_length_batch(sw, oracle):
    """C5-style mixture under one scoring scheme: tiny junction pairs, read-vs-read segments, long reads vs
    50-nt junctions, shuffled into one batch (every kernel instance and list class at once)"""
    from ciri_long_b200 import workloads as W
    b = W.concat_batches([W.junction_pairs(3000, seed=41), W.rolling_circle_pairs(40, seed=42, read_min=1500, read_max=5000),
                          W.long_query_short_ref_pairs(300, seed=43), W.square_pairs(40, 700, seed=44, params=(10, 4, 8, 2))],
                         shuffle_seed=45)
    rng = np.random.default_rng(46)
    check_batch(sw, oracle, b, sample=rng.choice(len(b), 500, replace=False))


def test_clip_vs_400kb_windows(sw, oracle):
    """S1 as it really is (find_bsj.py:191-215): clipped ends of 20-600 nt against 400-600 kb genomic windows that
    are views into one genome buffer (column-chunk tasks, bounded reverse look, fan-out of the per-list launches)"""
    import torch
    from ciri_long_b200 import workloads as W
    b = W.clip_window_pairs_torch(10, torch.device("cpu"), seed=61, genome_len=1 << 21, win_min=400000, win_max=600000)
    assert b.r_len.min() >= 400000
    check_batch(sw, oracle, b)
    rec, _ = run_batch(sw, b, flag=0)
    assert ((rec["status"] & 0xff) == 0).all()


@pytest.mark.parametrize("params", [(1, 1, 1, 1), (2, 2, 3, 1)])
def test_square_4096(sw, oracle, params):
    """C4's largest row at the find_bsj scoring too (four tiles of 1024 rows; 10/4/8/2 is covered by test_square_sweep
    up to 2048 and by test_scores_beyond_16_bit_comfort at 4096)"""
    from ciri_long_b200 import workloads as W
    b = W.square_pairs(3, 4096, params=params)
    check_batch(sw, oracle, b)
    check_batch(sw, oracle, b, flag=0)
