"""TEST INFRASTRUCTURE -- golden vectors for pairs whose CIGAR traceback leaves the band (ssw.c:642-673 then
reads direction bytes no band pass wrote: the reference's CIGAR is whatever that heap memory yields, or NULL
after "Trace back error").  Score and coordinates are well defined and are recorded here FROM THE UNMODIFIED
REFERENCE (oracle/_ref/libssw.so, built from /root/reference by oracle/Makefile); the CIGAR is recorded as
null.  The candidate pairs were found with the fuzz generator of tools/fuzz_gpu.py (family "mixed") run
against oracle/ssw_oracle.c, which reports that outcome as ORC_ERR_TRACEBACK.

    python oracle/make_golden_escape.py candidates.json     ->  tests/golden/band_escape.json
candidates.json: [[match, mismatch, gap_open, gap_extend], query codes, ref codes] per case."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    cands = json.load(open(sys.argv[1]))
    O.build(ref=True)
    ref, orc = O.RefLib(), O.Oracle()
    out = []
    for k, (params, q, r) in enumerate(cands):
        q, r = np.array(q, np.int8), np.array(r, np.int8)
        mat = O.make_mat(params[0], params[1])
        assert orc.align(q, r, mat, params[2], params[3], flag=1) is None, "the restatement must report the band escape"
        # flag 0 / flag 4 with a negative distance filter never enter banded_sw: the reference's own numbers
        e0 = ref.align(q, r, mat, params[2], params[3], flag=0)
        lib = ref.lib
        prof = lib.ssw_init(O._i8p(q), len(q), O._i8p(mat), 5, 2)
        res = lib.ssw_align(prof, O._i8p(r), len(r), params[2], params[3], 4, 0, -1, O.default_mask_len(len(q)))
        c = res.contents
        exp = dict(score=c.score1, score2=c.score2, ref_begin=c.ref_begin1, ref_end=c.ref_end1,
                   read_begin=c.read_begin1, read_end=c.read_end1, ref_end2=c.ref_end2, cigar=None)
        assert c.cigarLen == 0 and exp["score"] == e0["score"] and exp["ref_end"] == e0["ref_end"]
        lib.align_destroy(res)
        lib.init_destroy(prof)
        out.append(dict(name="band_escape_%02d" % k, params=list(params), query="".join("ACGTN"[x] for x in q),
                        ref="".join("ACGTN"[x] for x in r), expected=exp))
    with open(os.path.join(ROOT, "tests", "golden", "band_escape.json"), "w") as f:
        json.dump(dict(generator="oracle/make_golden_escape.py", source="unmodified reference ssw.c (flag=4, filterd=-1: "
                       "score + begin/end coordinates, no banded_sw)", cases=out), f, indent=0)
    print("wrote", len(out), "cases; sizes", [(len(c["query"]), len(c["ref"])) for c in out])


if __name__ == "__main__":
    main()
