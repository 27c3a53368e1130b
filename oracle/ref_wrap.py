"""TEST / BENCH INFRASTRUCTURE ONLY -- the per-call path of CIRI-long's `ssw_wrap.Aligner` restated, on top of the
UNMODIFIED reference library (oracle/_ref/libssw.so).

BASELINE.md section 3 asks for the CPU baseline "the way CIRI-long drives it": a new `Aligner(ref, ...)` per pair
(find_bsj.py:204-205, collapse.py:170-171), `.align(query)`, inside `multiprocessing.Pool` workers.  The
reference's own ssw_wrap.py cannot travel to the GPU box (nothing there may read /root/reference), so the work
that file does per call is restated here step by step -- same Python-level operations, same ctypes calls:

  * `_DNA_to_int_mat` (ssw_wrap.py:234-252): a c_int8 array filled by a per-base loop with a dict lookup inside
    try / except KeyError / finally -- the dominant cost for long references;
  * `set_mat` (ssw_wrap.py:146-159): a fresh 25-entry c_int8 matrix per Aligner;
  * `align` (ssw_wrap.py:174-230): ssw_init(score_size=2), mask length rule, ssw_align(flag=1), the score / length
    filter, a result object that copies the fields, init_destroy, align_destroy;
  * the CIGAR string (ssw_wrap.py:349-379) when `report_cigar`: two ctypes calls per op.

Nothing under ciri-long_b200/ imports this module.
"""
import ctypes as C

from . import oracle as O

_BASE = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4, 'a': 0, 'c': 1, 'g': 2, 't': 3, 'n': 4}


class RefAlignment(object):
    """fields of PyAlignRes (ssw_wrap.py:315-345)"""
    __slots__ = ("score", "ref_begin", "ref_end", "query_begin", "query_end", "score2", "ref_end2", "cigar_string", "cigar")


class RefAligner(object):
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = O.RefLib().lib
            lib.cigar_int_to_len.restype = C.c_int32
            lib.cigar_int_to_len.argtypes = [C.c_int32]
            lib.cigar_int_to_op.restype = C.c_char
            lib.cigar_int_to_op.argtypes = [C.c_int32]
            cls._lib = lib
        return cls._lib

    def __init__(self, ref_seq="", match=2, mismatch=2, gap_open=3, gap_extend=1, report_secondary=False, report_cigar=False):
        self.report_secondary, self.report_cigar = report_secondary, report_cigar
        self.gap_open, self.gap_extend = gap_open, gap_extend
        self.match, self.mismatch = match, mismatch
        x = -mismatch
        self.mat = (C.c_int8 * 25)(match, x, x, x, 0, x, match, x, x, 0, x, x, match, x, 0, x, x, x, match, 0, 0, 0, 0, 0, 0)
        if ref_seq:
            self.ref_len = len(ref_seq)
            self.ref_seq = self._to_codes(ref_seq, self.ref_len)
        else:
            self.ref_len, self.ref_seq = 0, ""

    @staticmethod
    def _to_codes(seq, n):
        arr = (C.c_int8 * n)()
        for i in range(n):                      # the reference's per-base loop, exception handler included
            try:
                v = _BASE[seq[i]]
            except KeyError:
                v = 4
            finally:
                arr[i] = v
        return arr

    def align(self, query_seq, min_score=0, min_len=0):
        lib = self.lib()
        qn = len(query_seq)
        q = self._to_codes(query_seq, qn)
        prof = lib.ssw_init(q, C.c_int32(qn), self.mat, 5, 2)
        mask_len = qn // 2 if qn > 30 else 15
        res = lib.ssw_align(prof, self.ref_seq, C.c_int32(self.ref_len), self.gap_open, self.gap_extend, 1, 0, 0, mask_len)
        if res and res.contents:
            score = res.contents.score1
            span = res.contents.read_end1 - res.contents.read_begin1 + 1
        else:
            score, span = -999999999999, -10000000000
        out = None
        if score >= min_score and span >= min_len:
            c = res.contents
            out = RefAlignment()
            out.score, out.ref_begin, out.ref_end = c.score1, c.ref_begin1, c.ref_end1
            out.query_begin, out.query_end = c.read_begin1, c.read_end1
            if self.report_secondary and c.score2 != 0:
                out.score2, out.ref_end2 = c.score2, c.ref_end2
            else:
                out.score2 = out.ref_end2 = None
            out.cigar = [int(c.cigar[i]) for i in range(c.cigarLen)]
            out.cigar_string = None
            if self.report_cigar and c.cigarLen > 0:
                s = ""
                if out.query_begin > 0:
                    s += "{}S".format(out.query_begin)
                for i in range(c.cigarLen):
                    s += "{}{}".format(lib.cigar_int_to_len(c.cigar[i]), lib.cigar_int_to_op(c.cigar[i]).decode())
                tail = qn - out.query_end - 1
                if tail != 0:
                    s += "{}S".format(tail)
                out.cigar_string = s
        lib.init_destroy(prof)
        if res:
            lib.align_destroy(res)
        return out
