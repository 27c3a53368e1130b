#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python oracle/make_golden.py

What it does
  1. builds oracle/_ref/libssw.so from the reference sources (make -C oracle ref);
  2. imports the reference's own Python wrapper (libs/striped_smith_waterman/ssw_wrap.py) from a scratch
     copy placed next to that libssw.so (the wrapper insists on loading "libssw.so" from its own
     directory, ssw_wrap.py:17, and /root/reference is read-only);
  3. runs the wrapper (Aligner(...).align(...), report_secondary/report_cigar on) AND the raw C ABI on
     every case and cross-checks the two;
  4. writes tests/golden/testfa.npz (the two sequences of tests/test.fa as int8 codes, the only data
     fixture the reference ships) and tests/golden/golden.json (inputs as strings + expected outputs).

The vectors pin: score, ref_begin/end, query_begin/end, score2, ref_end2, the raw CIGAR ops and the
wrapper's soft-clipped CIGAR string.
"""
import importlib.util
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF_ROOT = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
BASES = "ACGTN"


def load_reference_wrapper():
    tmp = tempfile.mkdtemp(prefix="sswref_")
    shutil.copy(os.path.join(REF_ROOT, "libs/striped_smith_waterman/ssw_wrap.py"), tmp)
    shutil.copy(O.REF_SO, os.path.join(tmp, "libssw.so"))
    spec = importlib.util.spec_from_file_location("ref_ssw_wrap", os.path.join(tmp, "ssw_wrap.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def to_str(codes):
    return "".join(BASES[c] for c in codes)


def channel(x, sub, ins, dele, maxrun, rng):
    out = []
    for b in x:
        r = rng.random()
        if r < dele:
            continue
        out.append((b + rng.integers(1, 4)) % 4 if r < dele + sub else b)
        if rng.random() < ins:
            out.extend(rng.integers(0, 4, size=rng.integers(1, maxrun + 1)).tolist())
    return np.array(out, dtype=np.int8)


def fuzz_cases(rng):
    """parameter sets x shapes x error models, overflow boundary, planted repeats."""
    cases = []
    params = [(1, 1, 1, 1), (10, 4, 8, 2), (2, 2, 3, 1), (2, 2, 2, 2)]
    k = 0
    for p in params:
        for shape in ("tiny", "c2", "square", "tallq"):
            for model in ("ont", "noisy", "random", "nrich"):
                for _ in range(3):
                    if shape == "tiny":
                        n, m = rng.integers(15, 61), rng.integers(15, 61)
                    elif shape == "c2":
                        n, m = rng.integers(500, 900), rng.integers(120, 420)
                    elif shape == "square":
                        n = rng.integers(150, 500); m = n + rng.integers(-15, 16)
                    else:
                        n, m = rng.integers(20, 60), rng.integers(100, 400)
                    r = rng.integers(0, 4, size=n).astype(np.int8)
                    if model == "random":
                        q = rng.integers(0, 4, size=m).astype(np.int8)
                    else:
                        st = rng.integers(0, max(1, n - min(m, n) + 1))
                        sub = r[st:st + m]
                        if model == "ont":
                            q = channel(sub, .05, .04, .04, 3, rng)
                        elif model == "noisy":
                            q = channel(sub, .12, .08, .08, 8, rng)
                        else:
                            q = channel(sub, .05, .04, .04, 3, rng)
                            q[rng.random(len(q)) < 0.10] = 4
                            r = r.copy(); r[rng.random(len(r)) < 0.05] = 4
                    if len(q) < 2:
                        continue
                    cases.append(dict(name="fuzz%03d_%s_%s" % (k, shape, model), params=p,
                                      ref=to_str(r), query=to_str(q)))
                    k += 1
    # overflow boundary: perfect matches whose score straddles 255 - bias
    for p, lens in (((1, 1, 1, 1), range(250, 259)), ((4, 2, 3, 1), range(60, 67)), ((10, 4, 8, 2), range(23, 28))):
        for L in lens:
            core = rng.integers(0, 4, size=L).astype(np.int8)
            r = np.concatenate([np.full(30, (core[0] + 1) % 4), core, np.full(30, (core[-1] + 1) % 4)]).astype(np.int8)
            q = np.concatenate([np.full(6, (core[0] + 2) % 4), core, np.full(4, (core[-1] + 2) % 4)]).astype(np.int8)
            cases.append(dict(name="ovf_%d_%d_%d_%d_L%d" % (p + (L,)), params=p, ref=to_str(r), query=to_str(q)))
    # planted repeats at distance maskLen-1, maskLen, maskLen+1 (both flavours)
    for p, L in (((1, 1, 1, 1), 60), ((1, 1, 1, 1), 300), ((10, 4, 8, 2), 40), ((10, 4, 8, 2), 90)):
        for d in (-1, 0, 1, 2):
            core = rng.integers(0, 4, size=L).astype(np.int8)
            q = core.copy()
            mask = O.default_mask_len(len(q))
            second = core.copy(); second[L // 2] = (second[L // 2] + 1) % 4     # slightly worse copy
            gap = mask + d
            filler = rng.integers(0, 4, size=max(gap - L, 0) + 0).astype(np.int8)
            r = np.concatenate([rng.integers(0, 4, size=25).astype(np.int8), core,
                                filler[:max(gap - L, 0)], second, rng.integers(0, 4, size=25).astype(np.int8)])
            cases.append(dict(name="rep_%d_%d_%d_%d_L%d_d%+d" % (p + (L, d)), params=p, ref=to_str(r), query=to_str(q)))
    # gap_open == gap_extend with match != mismatch: the reverse pass can jump over score1 without ever
    # equalling it (the stop rule of ssw.c:499 then fires late or never)
    for p in ((3, 2, 2, 2), (2, 1, 1, 1)):
        for t in range(30):
            n = int(rng.integers(300, 600))
            r = rng.integers(0, 4, size=n).astype(np.int8)
            q = channel(r, .04, .06, .10, 6, rng)
            cases.append(dict(name="revjump_%d_%d_%d_%d_%02d" % (p + (t,)), params=p, ref=to_str(r), query=to_str(q)))
    # hand-written vectors of SURVEY 8(c): G5 (score 0 / UB), G6 (N and unknown symbols), G7 (trap 3)
    cases.append(dict(name="G5_zero_score", params=(1, 1, 1, 1), ref="A" * 20, query="C" * 24))
    cases.append(dict(name="G6_n_and_unknown", params=(1, 1, 1, 1), ref="ACGTNNNNACGTACGTXX", query="ACGTACGTACGTACGTRY"))
    cases.append(dict(name="G7_leading_deletion", params=(10, 4, 8, 2), ref="GTGATTGCGTTTCTA", query="GGATATGACGACTA"))
    cases.append(dict(name="lowercase", params=(2, 2, 3, 1), ref="acgtacgtacgtTTGACCA", query="cgtacgtTTGAC"))
    # found by tools/fuzz_gpu.py (round 1): queries of more than 1024 rows (several strip tiles) whose reverse pass
    # meets score1 while strips ahead of the stop column already hold larger values ...
    for p in ((1, 1, 1, 1), (2, 1, 1, 1), (3, 2, 2, 2)):
        for t in range(10):
            n = int(rng.integers(1000, 1500)); m = int(rng.integers(1050, 1400))
            r = rng.integers(0, 4, size=n).astype(np.int8)
            L = min(m, n); st = int(rng.integers(0, n - L + 1))
            q = channel(r[st:st + L], .05, .05, .05, 4, rng)
            cases.append(dict(name="multitile_%d_%d_%d_%d_%02d" % (p + (t,)), params=p, ref=to_str(r), query=to_str(q)))
    # ... and periodic references, where several cells share the maximum and the first column must win even
    # though a later column is computed earlier (different query rows)
    for p in ((5, 5, 10, 3), (10, 4, 8, 2), (1, 1, 1, 1)):
        for t in range(10):
            n = int(rng.integers(300, 1400)); unit = rng.integers(0, 4, size=int(rng.integers(20, 200))).astype(np.int8)
            r = np.tile(unit, n // len(unit) + 1)[:n].astype(np.int8)
            m = int(rng.integers(100, 900)); L = min(m, n); st = int(rng.integers(0, n - L + 1))
            q = channel(r[st:st + L], .03, .03, .03, 3, rng)
            if len(q) < 2:
                continue
            cases.append(dict(name="periodic_%d_%d_%d_%d_%02d" % (p + (t,)), params=p, ref=to_str(r), query=to_str(q)))
    return cases


def run_case(wrap, reflib, ref_s, query_s, p):
    al = wrap.Aligner(ref_s, match=p[0], mismatch=p[1], gap_open=p[2], gap_extend=p[3],
                      report_secondary=True, report_cigar=True)
    pr = al.align(query_s)
    raw = reflib.align(O.encode(query_s), O.encode(ref_s), O.make_mat(p[0], p[1]), p[2], p[3])
    assert pr is not None and raw is not None
    assert (pr.score, pr.ref_begin, pr.ref_end, pr.query_begin, pr.query_end) == \
        (raw["score"], raw["ref_begin"], raw["ref_end"], raw["read_begin"], raw["read_end"])
    assert (pr.score2 or 0) == raw["score2"]
    cs = O.cigar_string(raw, len(query_s))
    assert pr.cigar_string == cs, (pr.cigar_string, cs)
    exp = {k: int(raw[k]) for k in O.FIELDS}
    exp["cigar"] = [int(c) for c in raw["cigar"]]
    exp["cigar_string"] = cs
    return exp


def main():
    O.build(ref=True)
    wrap = load_reference_wrapper()
    reflib = O.RefLib()
    os.makedirs(GOLD, exist_ok=True)

    with open(os.path.join(REF_ROOT, "tests/test.fa")) as f:
        f.readline(); seq1 = f.readline().strip(); f.readline(); seq2 = f.readline().strip()
    np.savez_compressed(os.path.join(GOLD, "testfa.npz"), seq1=O.encode(seq1), seq2=O.encode(seq2))

    out = dict(generator="oracle/make_golden.py",
               source="unmodified /root/reference/libs/striped_smith_waterman/ssw.c + ssw_wrap.py",
               testfa=[], cases=[])
    for name, r, q, p in (("G1", seq1, seq2, (1, 1, 1, 1)), ("G2", seq1, seq2, (10, 4, 8, 2)),
                          ("G3", seq2, seq1, (1, 1, 1, 1)), ("G4", seq2, seq1, (10, 4, 8, 2))):
        exp = run_case(wrap, reflib, r, q, p)
        out["testfa"].append(dict(name=name, ref="seq1" if r is seq1 else "seq2",
                                  query="seq1" if q is seq1 else "seq2", params=p, expected=exp))
        print(name, {k: exp[k] for k in O.FIELDS}, exp["cigar_string"][:60])

    rng = np.random.default_rng(20261017)
    for c in fuzz_cases(rng):
        c["expected"] = run_case(wrap, reflib, c["ref"], c["query"], c["params"])
        out["cases"].append(c)
    with open(os.path.join(GOLD, "golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote %d cases" % len(out["cases"]))


if __name__ == "__main__":
    main()
