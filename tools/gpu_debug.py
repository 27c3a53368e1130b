"""Ad-hoc GPU diagnostics: run golden cases + seeded batches through the CUDA path, print every mismatch."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
from oracle import oracle as O

orc = O.Oracle()
def check(b, flag=1, limit=5, sample=None):
    with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, b.match, b.mismatch, b.gap_open, b.gap_extend, flag=flag) as d:
        d.run(); rec, cig = d.fetch()
    mat = O.make_mat(b.match, b.mismatch)
    bad = 0; punt = 0
    for i in (range(len(b)) if sample is None else sample):
        e = orc.align(b.query(i), b.ref(i), mat, b.gap_open, b.gap_extend, flag=flag)
        r = rec[i]
        got = dict(score=int(r["score1"]), score2=int(r["score2"]), ref_begin=int(r["ref_begin1"]), ref_end=int(r["ref_end1"]),
                   read_begin=int(r["read_begin1"]), read_end=int(r["read_end1"]), ref_end2=int(r["ref_end2"]),
                   cigar=cig[r["cigar_off"]:r["cigar_off"]+r["cigar_len"]].tolist())
        st = int(r["status"])
        if st & 0xff:
            punt += 1
            if punt <= limit: print("  STATUS", i, hex(st), "m", b.q_len[i], "n", b.r_len[i], {k: got[k] for k in O.FIELDS}, "exp", {k: e[k] for k in O.FIELDS}, "word", e["word"])
        elif not O.same(got, e) or int(r["word"]) != e["word"]:
            bad += 1
            if bad <= limit: print("  MISMATCH", i, "m", b.q_len[i], "n", b.r_len[i], "got", {k: got[k] for k in O.FIELDS}, int(r["word"]), "exp", {k: e[k] for k in O.FIELDS}, e["word"], "cigar_eq", got["cigar"] == e["cigar"], "bw", e["band_width"])
    print("%-40s flag %d pairs %6d  mismatches %d  status!=0 %d" % (b.name, flag, len(b), bad, punt))
    return bad, punt

g = json.load(open("tests/golden/golden.json"))
byp = {}
for c in g["cases"]: byp.setdefault(tuple(c["params"]), []).append(c)
for p, cases in byp.items():
    b = W.from_lists([O.encode(c["query"]) for c in cases], [O.encode(c["ref"]) for c in cases], p, name="golden %s" % (p,))
    check(b)
if len(sys.argv) > 1 and sys.argv[1] == "quick": sys.exit(0)
for p in [(1,1,1,1),(10,4,8,2),(2,2,3,1),(2,2,2,2)]:
    check(W.bsj_refinement_pairs(200, seed=11, params=p))
    check(W.bsj_refinement_pairs(200, seed=11, params=p), flag=0)
check(W.rolling_circle_pairs(24, seed=12, read_min=800, read_max=3000))
for L in (64, 256, 1024, 2048):
    for p in [(1,1,1,1),(10,4,8,2)]:
        check(W.square_pairs(max(4, 2048//L), L, params=p))
check(W.junction_pairs(2000, seed=13))
check(W.overflow_boundary_pairs())
check(W.overflow_boundary_pairs(params=(2,2,2,2), lengths=range(118,134)))
