"""GPU parity tests of the throughput CIGAR pass (ciri-long_b200/csrc/ssw_tband.cu: one pair per lane, two rows
per s16x2 register) against the CPU oracle, through the C ABI.  The library picks that instance for batches of
at least SSW_CUDA_TBAND_MIN pairs (default 16384); the tests force it for small batches too, so that every
family -- tiny junction pairs, band doubling, the zeroed-neighbour rows of ssw.c:595-596, hand-over of wide
bands / score 0 / scores near the 16-bit range to the warp-per-pair instance -- is compared pair by pair."""
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


@pytest.fixture()
def tband(monkeypatch):
    monkeypatch.setenv("SSW_CUDA_TBAND_MIN", "0")


def compare_all(sw, oracle, b, flag=1):
    """every pair of the batch, field by field and op by op, against oracle.align_batch"""
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, b.match, b.mismatch, b.gap_open,
                        b.gap_extend, flag=flag) as d:
        d.run()
        rec, cig = d.fetch()
    ml = np.array([O.default_mask_len(int(x)) for x in b.q_len], dtype=np.int32)
    exp, ecig = oracle.align_batch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, O.make_mat(b.match, b.mismatch),
                                   b.gap_open, b.gap_extend, flag, ml)
    ok_status = exp["status"] == 0
    tb_err = exp["status"] == -3                       # the reference's "Trace back error" outcome
    assert (((rec["status"] & 0xff) == 0) == ok_status).all(), (b.name, np.nonzero(((rec["status"] & 0xff) == 0) != ok_status)[0][:5])
    assert (((rec["status"] & 0xff) == 1) == tb_err).all()
    for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2"):
        bad = np.nonzero(rec[k] != exp[k])[0]
        assert len(bad) == 0, (b.name, k, bad[:5], rec[k][bad[:5]], exp[k][bad[:5]])
    bad = np.nonzero((rec["cigar_len"] != exp["cigar_len"]) & ok_status)[0]
    assert len(bad) == 0, (b.name, "cigar_len", bad[:5], rec["cigar_len"][bad[:5]], exp["cigar_len"][bad[:5]])
    for i in np.nonzero(ok_status)[0]:
        g = cig[rec["cigar_off"][i]:rec["cigar_off"][i] + rec["cigar_len"][i]]
        e = ecig[exp["cigar_off"][i]:exp["cigar_off"][i] + exp["cigar_len"][i]]
        assert (g == e).all(), (b.name, int(i), g.tolist()[:12], e.tolist()[:12], int(exp["band_width"][i]))
    return rec, exp


@pytest.mark.parametrize("params", [(1, 1, 1, 1), (10, 4, 8, 2), (2, 2, 3, 1), (2, 2, 2, 2)])
def test_tband_bsj_pairs(sw, oracle, tband, params):
    from ciri_long_b200 import workloads as W
    compare_all(sw, oracle, W.bsj_refinement_pairs(1500, seed=31, params=params))


def test_tband_tiny_junction_pairs(sw, oracle, tband):
    from ciri_long_b200 import workloads as W
    rec, exp = compare_all(sw, oracle, W.junction_pairs(6000, seed=32))
    assert (exp["band_width"] > 1).any()


def test_tband_rolling_circle(sw, oracle, tband):
    from ciri_long_b200 import workloads as W
    rec, exp = compare_all(sw, oracle, W.rolling_circle_pairs(48, seed=33, read_min=800, read_max=4000))
    assert exp["band_width"].max() >= 32           # several doubling passes happened


@pytest.mark.parametrize("params", [(1, 1, 1, 1), (10, 4, 8, 2)])
def test_tband_squares_and_mixed_sizes(sw, oracle, tband, params):
    """one batch with pairs of very different size: lock step across unequal lanes"""
    from ciri_long_b200 import workloads as W
    parts = [W.square_pairs(40, L, params=params) for L in (16, 64, 200, 700)]
    parts.append(W.junction_pairs(300, seed=34, params=params))
    compare_all(sw, oracle, W.concat_batches(parts, shuffle_seed=5))


def test_tband_hand_over_cases(sw, oracle, tband):
    """pairs the lane kernel must hand to the warp-per-pair instance: score 0, long indels (bands wider than
    126 diagonals), scores near the 16-bit range; plus degenerate shapes"""
    from ciri_long_b200 import workloads as W
    rng = np.random.default_rng(35)
    qs, rs = [], []
    for k in range(40):                                           # long deletions / insertions: wide bands
        core = rng.integers(0, 4, 600).astype(np.int8)
        gap = int(rng.integers(100, 400))
        a = np.concatenate([core[:300], rng.integers(0, 4, gap).astype(np.int8), core[300:]])
        if k % 2:
            qs.append(core); rs.append(a)
        else:
            qs.append(a); rs.append(core)
    qs += [np.zeros(24, np.int8) + 1, np.array([0], np.int8), np.array([4, 4, 4, 4], np.int8), np.arange(200, dtype=np.int8) % 4]
    rs += [np.zeros(20, np.int8), np.array([0], np.int8), np.array([0, 1, 2, 3], np.int8), np.arange(37, dtype=np.int8) % 4]
    for p in [(1, 1, 1, 1), (10, 4, 8, 2)]:
        compare_all(sw, oracle, W.from_lists(qs, rs, p, name="hand-over"))
    big = rng.integers(0, 4, 3400).astype(np.int8)                # perfect 3.4 kb match at 10 per base: saturates int16
    compare_all(sw, oracle, W.from_lists([big, big[:3190]], [big, big[:3190]], (10, 4, 8, 2), name="near-int16"))


def test_tband_golden_cases(sw, golden, tband):
    """every committed golden vector (generated from the unmodified reference) through the lane kernel"""
    from ciri_long_b200 import workloads as W
    by_params = {}
    for c in golden["cases"]:
        by_params.setdefault(tuple(c["params"]), []).append(c)
    for p, cases in by_params.items():
        b = W.from_lists([O.encode(c["query"]) for c in cases], [O.encode(c["ref"]) for c in cases], p)
        with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, *p, flag=1) as d:
            d.run()
            rec, cig = d.fetch()
        for i, c in enumerate(cases):
            e, r = c["expected"], rec[i]
            assert (r["status"] & 0xff) == 0, (c["name"], int(r["status"]))
            assert cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]].tolist() == e["cigar"], c["name"]


def test_tband_default_threshold_and_repeat(sw, oracle):
    """a batch above the default threshold takes the lane kernel on its own; two runs of the same resident batch
    agree, and so does the warp-per-pair instance"""
    import os
    from ciri_long_b200 import workloads as W
    b = W.bsj_refinement_pairs(20000, seed=36)
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1) as d:
        d.run(); r1, c1 = d.fetch()
        d.run(); r2, c2 = d.fetch()
    os.environ["SSW_CUDA_TBAND_MIN"] = "1000000000"
    try:
        with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1) as d:
            d.run(); r3, c3 = d.fetch()
    finally:
        del os.environ["SSW_CUDA_TBAND_MIN"]
    for ra, ca in ((r2, c2), (r3, c3)):
        for k in ("score1", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "score2", "ref_end2", "cigar_len", "status"):
            assert (r1[k] == ra[k]).all(), k
        o1, oa = np.argsort(r1["cigar_off"], kind="stable"), np.argsort(ra["cigar_off"], kind="stable")
        # CIGAR windows are handed out in completion order: compare pair by pair
        for i in range(0, len(b), 97):
            assert (c1[r1["cigar_off"][i]:r1["cigar_off"][i] + r1["cigar_len"][i]] ==
                    ca[ra["cigar_off"][i]:ra["cigar_off"][i] + ra["cigar_len"][i]]).all(), i


@pytest.mark.parametrize("lane_kernel", [False, True])
def test_band_escape_returns_coordinates_and_empty_cigar(sw, monkeypatch, lane_kernel):
    """tests/golden/band_escape.json (score and coordinates from the unmodified reference): the traceback of these
    pairs leaves the band, where the reference's CIGAR is undefined.  Both CIGAR instances, the legacy per-call
    ABI and the Python wrapper return the exact coordinates with status 1 and an empty CIGAR."""
    import json, os
    GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    from ciri_long_b200 import workloads as W
    monkeypatch.setenv("SSW_CUDA_TBAND_MIN", "0" if lane_kernel else "1000000000")
    g = json.load(open(os.path.join(GOLDEN_DIR, "band_escape.json")))
    for c in g["cases"]:
        p, e = c["params"], c["expected"]
        b = W.from_lists([O.encode(c["query"])], [O.encode(c["ref"])], tuple(p))
        with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, *p, flag=1) as d:
            d.run()
            rec, cig = d.fetch()
        r = rec[0]
        assert (int(r["status"]) & 0xff) == 1 and int(r["cigar_len"]) == 0, (c["name"], int(r["status"]))
        got = tuple(int(r[k]) for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2"))
        assert got == tuple(e[k] for k in O.FIELDS), (c["name"], got)
        al = sw.Aligner(c["ref"], p[0], p[1], p[2], p[3], report_secondary=True, report_cigar=True)
        res = al.align(c["query"])
        assert res is not None and res.cigar_string is None
        assert (res.score, res.ref_begin, res.ref_end, res.query_begin, res.query_end) == \
            (e["score"], e["ref_begin"], e["ref_end"], e["read_begin"], e["read_end"])
        res2 = sw.align_pairs([c["ref"]], [c["query"]], *p, report_cigar=True)[0]
        assert res2 is not None and res2.cigar_string is None and res2.ref_begin == e["ref_begin"]


def test_packed_input_matches_unpacked(sw):
    """ssw_batch_create_packed / ssw_align_batch_multi_packed (two bases per byte, offsets in bases, odd offsets and
    lengths, N bases) give the records and CIGARs of the one-code-per-byte calls"""
    from ciri_long_b200 import workloads as W
    parts = [W.bsj_refinement_pairs(700, seed=41, params=(10, 4, 8, 2)), W.junction_pairs(900, seed=42), W.square_pairs(60, 333, params=(10, 4, 8, 2))]
    b = W.concat_batches(parts, shuffle_seed=3)
    packed = sw.pack_dna4(b.seqs)
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 10, 4, 8, 2) as d:
        d.run(); r1, c1 = d.fetch()
    with sw.DeviceBatch(packed, b.q_off, b.q_len, b.r_off, b.r_len, 10, 4, 8, 2, packed_bases=len(b.seqs)) as d:
        d.run(); r2, c2 = d.fetch()
    r3, c3 = sw.align_arrays(packed, b.q_off, b.q_len, b.r_off, b.r_len, 10, 4, 8, 2, packed_bases=len(b.seqs))
    for ra, ca in ((r2, c2), (r3, c3)):
        for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len", "status", "word"):
            assert (r1[k] == ra[k]).all(), k
        for i in range(len(b)):
            assert (c1[r1["cigar_off"][i]:r1["cigar_off"][i] + r1["cigar_len"][i]] == ca[ra["cigar_off"][i]:ra["cigar_off"][i] + ra["cigar_len"][i]]).all()


def test_multi_device_call_matches_single_device(sw, monkeypatch):
    """ssw_align_batch_multi over all devices of the box (chunks pulled from a shared queue by one host thread per
    device) = the single-device call, pair by pair; with one device it degenerates to several chunks in flight"""
    from ciri_long_b200 import workloads as W
    ndev = sw.Aligner.libssw.ssw_cuda_device_count()
    monkeypatch.setenv("SSW_CUDA_CHUNK", "3000")                   # many chunks, so that every device gets several
    monkeypatch.setenv("SSW_CUDA_TBAND_MIN", "1024")
    b = W.concat_batches([W.bsj_refinement_pairs(9000, seed=43), W.junction_pairs(12000, seed=44, params=(1, 1, 1, 1))], shuffle_seed=9)
    b = W.repack(b)
    r1, c1 = sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, device=0)
    r2, c2 = sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, devices=list(range(ndev)) if ndev > 1 else [0])
    for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len", "status", "word"):
        assert (r1[k] == r2[k]).all(), k
    for i in range(0, len(b), 7):
        assert (c1[r1["cigar_off"][i]:r1["cigar_off"][i] + r1["cigar_len"][i]] == c2[r2["cigar_off"][i]:r2["cigar_off"][i] + r2["cigar_len"][i]]).all()
    assert len(c1) == len(c2) == int(r1["cigar_len"].sum())


@pytest.mark.parametrize("lane_kernel", [False, True])
def test_band_wider_than_the_scratch_is_rerun(sw, oracle, monkeypatch, lane_kernel):
    """a 700-900-nt deletion joined by two strong flanks (10/4/8/2 pays for it): the first band is already ~1700
    diagonals wide, the direction matrix does not fit the per-warp scratch and the pair is re-run from
    ssw_batch_fetch with a scratch sized for it -- same CIGAR as the reference's single malloc'ed matrix"""
    from ciri_long_b200 import workloads as W
    monkeypatch.setenv("SSW_CUDA_TBAND_MIN", "0" if lane_kernel else "1000000000")
    rng = np.random.default_rng(51)
    qs, rs = [], []
    for k in range(6):
        a, c = rng.integers(0, 4, 320).astype(np.int8), rng.integers(0, 4, 320).astype(np.int8)
        junk = rng.integers(0, 4, int(rng.integers(700, 900))).astype(np.int8)
        read, ref = np.concatenate([a, c]), np.concatenate([a, junk, c])
        if k % 2:
            read, ref = ref, read                                  # the same as a long insertion
        qs.append(read); rs.append(ref)
    rec, exp = compare_all(sw, oracle, W.from_lists(qs, rs, (10, 4, 8, 2), name="huge-gap"))
    assert (exp["band_width"] > 512).any()
