"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU checkers.

* ``Oracle``  : oracle/liboracle.so, the scalar restatement in ssw_oracle.c.
* ``RefLib``  : oracle/_ref/libssw.so, the UNMODIFIED reference ssw.c compiled by ``make -C oracle ref``
                (present only where /root/reference was available at build time, or shipped as a
                prebuilt binary to the GPU box).

Both expose ``align(read, ref, mat, go, ge, flag, mask_len) -> dict`` with the seven s_align fields
(ssw.h:42-52) and the CIGAR as a list of BAM-style uint32 ops, so tests can compare them (and the CUDA
path) field by field.  Nothing under ciri-long_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libssw.so")
REF_SRC = "/root/reference/libs/striped_smith_waterman"

FIELDS = ("score", "score2", "ref_begin", "ref_end", "read_begin", "read_end", "ref_end2")

_LUT = np.full(256, 4, dtype=np.int8)
for _i, _ch in enumerate("ACGTN"):
    _LUT[ord(_ch)] = _i
    _LUT[ord(_ch.lower())] = _i


def encode(seq):
    """ASCII -> {0..4}; anything that is not ACGTN (either case) becomes 4 (ssw_wrap.py:234-252)."""
    if isinstance(seq, str):
        seq = seq.encode("latin-1", "replace")
    return _LUT[np.frombuffer(seq, dtype=np.uint8)].copy()


def make_mat(match, mismatch):
    """5x5 matrix with a zero N row/column (ssw_wrap.py:146-159)."""
    m = np.full((5, 5), -mismatch, dtype=np.int8)
    np.fill_diagonal(m, match)
    m[4, :] = 0
    m[:, 4] = 0
    return m.reshape(-1).copy()


def default_mask_len(query_len):
    """ssw_wrap.py:196-199"""
    return query_len // 2 if query_len > 30 else 15


def build(ref=True):
    """Compile the checkers (never the product)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def cigar_ops(cigar):
    return [(int(c) >> 4, "MIDNSHP=X"[int(c) & 0xF] if (int(c) & 0xF) < 9 else "M") for c in cigar]


def cigar_string(res, query_len):
    """The wrapper's SAM-like string with soft clips (ssw_wrap.py:349-379)."""
    s = ""
    if res["read_begin"] > 0:
        s += "%dS" % res["read_begin"]
    for ln, op in cigar_ops(res["cigar"]):
        s += "%d%s" % (ln, op)
    tail = query_len - res["read_end"] - 1
    if tail != 0:
        s += "%dS" % tail
    return s


class _OrcResult(C.Structure):
    _fields_ = [("score1", C.c_uint16), ("score2", C.c_uint16),
                ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32),
                ("read_begin1", C.c_int32), ("read_end1", C.c_int32), ("ref_end2", C.c_int32),
                ("cigar", C.POINTER(C.c_uint32)), ("cigarLen", C.c_int32),
                ("word", C.c_int32), ("band_width", C.c_int32)]


class _OrcFlat(C.Structure):
    _fields_ = [("status", C.c_int32), ("word", C.c_int32), ("band_width", C.c_int32),
                ("score1", C.c_int32), ("score2", C.c_int32),
                ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32),
                ("read_begin1", C.c_int32), ("read_end1", C.c_int32), ("ref_end2", C.c_int32),
                ("cigar_off", C.c_int64), ("cigar_len", C.c_int32), ("_pad", C.c_int32)]


FLAT_DTYPE = np.dtype([("status", "<i4"), ("word", "<i4"), ("band_width", "<i4"),
                       ("score1", "<i4"), ("score2", "<i4"),
                       ("ref_begin1", "<i4"), ("ref_end1", "<i4"),
                       ("read_begin1", "<i4"), ("read_end1", "<i4"), ("ref_end2", "<i4"),
                       ("cigar_off", "<i8"), ("cigar_len", "<i4"), ("_pad", "<i4")])


def _i8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int8))


class Oracle:
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        self.lib.orc_align.restype = C.c_int
        self.lib.orc_align.argtypes = [C.POINTER(C.c_int8), C.c_int32, C.POINTER(C.c_int8), C.c_int32,
                                       C.POINTER(C.c_int8), C.c_int32, C.c_int8,
                                       C.c_uint8, C.c_uint8, C.c_uint8, C.c_uint16, C.c_int32, C.c_int32,
                                       C.POINTER(_OrcResult)]
        self.lib.orc_result_free.argtypes = [C.POINTER(_OrcResult)]
        self.lib.orc_align_batch.restype = C.c_int
        self.lib.orc_align_batch.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int32, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]

    def align(self, read, ref, mat, go, ge, flag=1, mask_len=None, score_size=2, n=5):
        read = np.ascontiguousarray(read, dtype=np.int8)
        ref = np.ascontiguousarray(ref, dtype=np.int8)
        mat = np.ascontiguousarray(mat, dtype=np.int8)
        if mask_len is None:
            mask_len = default_mask_len(len(read))
        r = _OrcResult()
        st = self.lib.orc_align(_i8p(read), len(read), _i8p(ref), len(ref), _i8p(mat), n, score_size,
                                go, ge, flag, 0, 0, mask_len, C.byref(r))
        if st != 0:
            self.lib.orc_result_free(C.byref(r))
            return None
        out = dict(score=r.score1, score2=r.score2, ref_begin=r.ref_begin1, ref_end=r.ref_end1,
                   read_begin=r.read_begin1, read_end=r.read_end1, ref_end2=r.ref_end2,
                   cigar=[int(r.cigar[i]) for i in range(r.cigarLen)], word=r.word, band_width=r.band_width)
        self.lib.orc_result_free(C.byref(r))
        return out

    def align_batch(self, seqs, q_off, q_len, r_off, r_len, mat, go, ge, flag, mask_len, cigar_cap=None):
        """Struct-of-arrays batch (single thread).  Returns (records[FLAT_DTYPE], cigar_buf[:used])."""
        n_pairs = len(q_len)
        seqs = np.ascontiguousarray(seqs, dtype=np.int8)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        r_off = np.ascontiguousarray(r_off, dtype=np.int64)
        q_len = np.ascontiguousarray(q_len, dtype=np.int32)
        r_len = np.ascontiguousarray(r_len, dtype=np.int32)
        mask_len = np.ascontiguousarray(mask_len, dtype=np.int32)
        mat = np.ascontiguousarray(mat, dtype=np.int8)
        if cigar_cap is None:
            cigar_cap = int(2 * (q_len.astype(np.int64) + r_len).sum() + 16)
        out = np.zeros(n_pairs, dtype=FLAT_DTYPE)
        cig = np.zeros(cigar_cap, dtype=np.uint32)
        used = C.c_int64(0)
        st = self.lib.orc_align_batch(n_pairs, seqs.ctypes.data, q_off.ctypes.data, q_len.ctypes.data,
                                      r_off.ctypes.data, r_len.ctypes.data, mat.ctypes.data, 5, go, ge, flag,
                                      mask_len.ctypes.data, out.ctypes.data, cig.ctypes.data, cigar_cap,
                                      C.byref(used))
        if st != 0:
            raise RuntimeError("orc_align_batch failed: %d" % st)
        return out, cig[:used.value]


class _SAlign(C.Structure):
    """s_align, ssw.h:42-52"""
    _fields_ = [("score1", C.c_uint16), ("score2", C.c_uint16),
                ("ref_begin1", C.c_int32), ("ref_end1", C.c_int32),
                ("read_begin1", C.c_int32), ("read_end1", C.c_int32), ("ref_end2", C.c_int32),
                ("cigar", C.POINTER(C.c_uint32)), ("cigarLen", C.c_int32)]


class RefLib:
    """The unmodified reference libssw.so driven exactly like ssw_wrap.py:187-227 (minus the Python
    per-base encode loop)."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = C.CDLL(path)
        lib.ssw_init.restype = C.c_void_p
        lib.ssw_init.argtypes = [C.POINTER(C.c_int8), C.c_int32, C.POINTER(C.c_int8), C.c_int32, C.c_int8]
        lib.init_destroy.restype = None
        lib.init_destroy.argtypes = [C.c_void_p]
        lib.ssw_align.restype = C.POINTER(_SAlign)
        lib.ssw_align.argtypes = [C.c_void_p, C.POINTER(C.c_int8), C.c_int32, C.c_uint8, C.c_uint8, C.c_uint8,
                                  C.c_uint16, C.c_int32, C.c_int32]
        lib.align_destroy.restype = None
        lib.align_destroy.argtypes = [C.POINTER(_SAlign)]
        self.lib = lib

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def align(self, read, ref, mat, go, ge, flag=1, mask_len=None, score_size=2, n=5):
        read = np.ascontiguousarray(read, dtype=np.int8)
        ref = np.ascontiguousarray(ref, dtype=np.int8)
        mat = np.ascontiguousarray(mat, dtype=np.int8)
        if mask_len is None:
            mask_len = default_mask_len(len(read))
        prof = self.lib.ssw_init(_i8p(read), len(read), _i8p(mat), n, score_size)
        res = self.lib.ssw_align(prof, _i8p(ref), len(ref), go, ge, flag, 0, 0, mask_len)
        out = None
        if res:
            r = res.contents
            out = dict(score=r.score1, score2=r.score2, ref_begin=r.ref_begin1, ref_end=r.ref_end1,
                       read_begin=r.read_begin1, read_end=r.read_end1, ref_end2=r.ref_end2,
                       cigar=[int(r.cigar[i]) for i in range(r.cigarLen)])
            self.lib.align_destroy(res)
        self.lib.init_destroy(prof)
        return out


def same(a, b, with_cigar=True):
    if a is None or b is None:
        return a is b
    for k in FIELDS:
        if a[k] != b[k]:
            return False
    return (not with_cigar) or list(a["cigar"]) == list(b["cigar"])


def edit_distance(x, y):
    """Unit-cost global edit distance of two byte strings (oracle/edit_oracle.c; CIRI_long/utils.py:153-159)."""
    lib = C.CDLL(ORACLE_SO)
    lib.orc_edit_distance.restype = C.c_int32
    lib.orc_edit_distance.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
    if isinstance(x, str):
        x = x.encode("latin-1")
    if isinstance(y, str):
        y = y.encode("latin-1")
    return int(lib.orc_edit_distance(bytes(x), len(x), bytes(y), len(y)))
