"""CPU tests of the checker itself: the scalar restatement (oracle/ssw_oracle.c) must reproduce the
golden vectors generated from the unmodified reference (oracle/make_golden.py) and, where the reference
build is present, the live reference on seeded fuzz.  No GPU, no product code."""
import numpy as np
import pytest

from oracle import oracle as O


def _check(res, exp, query_len):
    assert res is not None
    for k in O.FIELDS:
        assert res[k] == exp[k], (k, res[k], exp[k])
    assert list(res["cigar"]) == list(exp["cigar"])
    assert O.cigar_string(res, query_len) == exp["cigar_string"]


def test_golden_cases(oracle, golden):
    assert len(golden["cases"]) > 200
    for c in golden["cases"]:
        p = c["params"]
        q, r = O.encode(c["query"]), O.encode(c["ref"])
        res = oracle.align(q, r, O.make_mat(p[0], p[1]), p[2], p[3])
        _check(res, c["expected"], len(q))


def test_golden_testfa(oracle, golden):
    """tests/test.fa of the reference, both orientations x both CIRI-long parameter sets (G1-G4)."""
    for c in golden["testfa"]:
        p = c["params"]
        q, r = golden["seqs"][c["query"]], golden["seqs"][c["ref"]]
        res = oracle.align(q, r, O.make_mat(p[0], p[1]), p[2], p[3])
        _check(res, c["expected"], len(q))
        assert res["word"] == 1          # all four take the int8 -> int16 re-run


def test_encode_matches_wrapper_rules():
    assert O.encode("ACGTNacgtnXR-").tolist() == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 4, 4, 4]
    m = O.make_mat(10, 4).reshape(5, 5)
    assert m[0, 0] == 10 and m[0, 1] == -4 and (m[4] == 0).all() and (m[:, 4] == 0).all()
    assert O.default_mask_len(30) == 15 and O.default_mask_len(31) == 15 and O.default_mask_len(500) == 250


def test_score_only_flag0(oracle, golden):
    c = golden["cases"][5]
    p = c["params"]
    q, r = O.encode(c["query"]), O.encode(c["ref"])
    res = oracle.align(q, r, O.make_mat(p[0], p[1]), p[2], p[3], flag=0)
    assert res["ref_begin"] == -1 and res["read_begin"] == -1 and res["cigar"] == []
    for k in ("score", "score2", "ref_end", "read_end", "ref_end2"):
        assert res[k] == c["expected"][k]


def test_against_live_reference(oracle, reflib):
    """Seeded fuzz against oracle/_ref/libssw.so (skipped on boxes without the reference build)."""
    from ciri_long_b200 import workloads as W
    batches = [W.bsj_refinement_pairs(40, seed=7), W.rolling_circle_pairs(12, seed=8, read_min=600, read_max=1500),
               W.square_pairs(20, 128, seed=9, params=(10, 4, 8, 2)), W.junction_pairs(300, seed=10),
               W.overflow_boundary_pairs(), W.overflow_boundary_pairs(params=(2, 2, 2, 2), lengths=range(120, 132))]
    n = 0
    for b in batches:
        mat = O.make_mat(b.match, b.mismatch)
        for i in range(len(b)):
            a = oracle.align(b.query(i), b.ref(i), mat, b.gap_open, b.gap_extend)
            r = reflib.align(b.query(i), b.ref(i), mat, b.gap_open, b.gap_extend)
            assert O.same(a, r), (b.name, i)
            n += 1
    assert n > 400


def test_batch_driver_matches_single(oracle):
    from ciri_long_b200 import workloads as W
    b = W.junction_pairs(64, seed=3)
    mat = O.make_mat(b.match, b.mismatch)
    mask = np.array([O.default_mask_len(int(x)) for x in b.q_len], dtype=np.int32)
    rec, cig = oracle.align_batch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, mat, b.gap_open, b.gap_extend, 1, mask)
    for i in range(len(b)):
        a = oracle.align(b.query(i), b.ref(i), mat, b.gap_open, b.gap_extend)
        assert rec["score1"][i] == a["score"] and rec["ref_begin1"][i] == a["ref_begin"]
        assert rec["read_end1"][i] == a["read_end"] and rec["ref_end2"][i] == a["ref_end2"]
        assert cig[rec["cigar_off"][i]:rec["cigar_off"][i] + rec["cigar_len"][i]].tolist() == a["cigar"]


def _py_edit_distance(x, y):
    """independent restatement (full matrix, pure Python) used only to cross-check the C checker"""
    D = [[0] * (len(y) + 1) for _ in range(len(x) + 1)]
    for i in range(len(x) + 1):
        D[i][0] = i
    for j in range(len(y) + 1):
        D[0][j] = j
    for i in range(1, len(x) + 1):
        for j in range(1, len(y) + 1):
            D[i][j] = min(D[i - 1][j] + 1, D[i][j - 1] + 1, D[i - 1][j - 1] + (x[i - 1] != y[j - 1]))
    return D[len(x)][len(y)]


def test_edit_distance_oracle_known_answers_and_properties():
    """oracle/edit_oracle.c (utils.py:153-159): published known answers, an independent restatement, metric
    properties.  edlib / python-Levenshtein themselves are absent from this image."""
    O.build(ref=False)
    for x, y, d in (("kitten", "sitting", 3), ("flaw", "lawn", 2), ("intention", "execution", 5),
                    ("GATTACA", "GCATGCU", 4), ("sunday", "saturday", 3), ("", "", 0), ("", "abc", 3), ("abc", "", 3),
                    ("ACGT", "ACGT", 0), ("acgt", "ACGT", 4)):
        assert O.edit_distance(x, y) == d
    rng = np.random.default_rng(3)
    seqs = ["".join(np.array(list("ACGTN"))[rng.integers(0, 5, int(rng.integers(0, 70)))]) for _ in range(40)]
    for a in seqs[:20]:
        for b in seqs[20:]:
            d = O.edit_distance(a, b)
            assert d == _py_edit_distance(a, b)
            assert d == O.edit_distance(b, a)
            assert abs(len(a) - len(b)) <= d <= max(len(a), len(b))
    a, b, c = seqs[1], seqs[2], seqs[3]
    assert O.edit_distance(a, c) <= O.edit_distance(a, b) + O.edit_distance(b, c)


def test_band_escape_goldens_are_reported_by_the_restatement(oracle):
    """tests/golden/band_escape.json: pairs whose traceback leaves the band.  The reference's CIGAR is undefined
    there (it reads direction bytes it never wrote, ssw.c:642-673); the restatement reports the outcome, and its
    score / coordinates equal the reference's own (recorded with flag=4, filterd=-1)."""
    import json, os
    GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = json.load(open(os.path.join(GOLDEN_DIR, "band_escape.json")))
    assert len(g["cases"]) >= 3
    for c in g["cases"]:
        q, r, p = O.encode(c["query"]), O.encode(c["ref"]), c["params"]
        assert oracle.align(q, r, O.make_mat(p[0], p[1]), p[2], p[3], flag=1) is None
        e = oracle.align(q, r, O.make_mat(p[0], p[1]), p[2], p[3], flag=0)
        assert (e["score"], e["ref_end"], e["read_end"], e["score2"], e["ref_end2"]) == \
            tuple(c["expected"][k] for k in ("score", "ref_end", "read_end", "score2", "ref_end2"))
