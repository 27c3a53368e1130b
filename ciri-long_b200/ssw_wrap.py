"""
ssw_wrap -- drop-in mirror of CIRI-long's ``libs/striped_smith_waterman/ssw_wrap.py`` on top of
``libssw_cuda.so`` (hand-written sm_100a kernels, no CPU path).

Same public surface as the reference module:

* ``Aligner(ref_seq, match, mismatch, gap_open, gap_extend, report_secondary, report_cigar)`` with
  ``set_gap / set_mat / set_ref / align(query_seq, min_score, min_len)``        (ssw_wrap.py:102-230)
* ``PyAlignRes`` with ``score, ref_begin, ref_end, query_begin, query_end, score2, ref_end2,
  cigar_string``                                                                 (ssw_wrap.py:315-379)
* ``CAlignRes`` (ctypes mirror of ``s_align``)                                   (ssw_wrap.py:29-37)

plus the batched entry points the CIRI-long call sites are moved to:

* ``Aligner.align_batch(queries)``      one reference, many queries  (collapse.py:212-265, 373-387)
* ``align_pairs(refs, queries, ...)``   many references              (find_bsj.py:204-215, collapse.py:170)
* ``DeviceBatch``                       struct-of-arrays access to ``ssw_batch_*`` (bench / pipelines)

The library is loaded at class-body time exactly like the reference (ssw_wrap.py:54,278): importing this
module without a built ``libssw_cuda.so`` raises, and every alignment call needs a CUDA device.
"""
import os
from ctypes import (POINTER, Structure, byref, c_char, c_char_p, c_double, c_int, c_int8, c_int32,
                    c_int64, c_uint8, c_uint16, c_uint32, c_void_p, cdll)

import numpy as np

_so_file_name = os.path.dirname(os.path.abspath(__file__)) + '/' + "libssw_cuda.so"

_ENC = np.full(256, 4, dtype=np.int8)
for _i, _b in enumerate("ACGTN"):
    _ENC[ord(_b)] = _i
    _ENC[ord(_b.lower())] = _i


_ENC_TABLE = bytes(int(v) for v in _ENC)


def encode_dna(seq):
    """ASCII -> {0..4} (A C G T N, either case, anything else 4): ssw_wrap.py:234-252 without the
    per-base Python loop."""
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.int8)
    if isinstance(seq, str):
        seq = seq.encode("latin-1", "replace")
    return np.frombuffer(bytes(seq).translate(_ENC_TABLE), dtype=np.int8)


class CAlignRes(Structure):
    """ctypes mirror of s_align (ssw.h:42-52), as in the reference wrapper."""
    _fields_ = [('score', c_uint16),
                ('score2', c_uint16),
                ('ref_begin', c_int32),
                ('ref_end', c_int32),
                ('query_begin', c_int32),
                ('query_end', c_int32),
                ('ref_end2', c_int32),
                ('cigar', POINTER(c_uint32)),
                ('cigarLen', c_int32)]


class SSWResult(Structure):
    """ssw_result of include/ssw_cuda.h"""
    _fields_ = [('score1', c_int32), ('score2', c_int32), ('ref_begin1', c_int32), ('ref_end1', c_int32),
                ('read_begin1', c_int32), ('read_end1', c_int32), ('ref_end2', c_int32), ('cigar_len', c_int32),
                ('cigar_off', c_int64), ('status', c_int32), ('word', c_int32)]


RESULT_DTYPE = np.dtype([('score1', '<i4'), ('score2', '<i4'), ('ref_begin1', '<i4'), ('ref_end1', '<i4'),
                         ('read_begin1', '<i4'), ('read_end1', '<i4'), ('ref_end2', '<i4'), ('cigar_len', '<i4'),
                         ('cigar_off', '<i8'), ('status', '<i4'), ('word', '<i4')])


class SSWScoring(Structure):
    _fields_ = [('mat', c_int8 * 25), ('gap_open', c_uint8), ('gap_extend', c_uint8), ('flag', c_uint8),
                ('_pad', c_uint8), ('filters', c_uint16), ('_pad2', c_uint16), ('filterd', c_int32)]


def _load():
    lib = cdll.LoadLibrary(_so_file_name)
    lib.ssw_init.restype = c_void_p
    lib.ssw_init.argtypes = [POINTER(c_int8), c_int32, POINTER(c_int8), c_int32, c_int8]
    lib.init_destroy.restype = None
    lib.init_destroy.argtypes = [c_void_p]
    lib.ssw_align.restype = POINTER(CAlignRes)
    lib.ssw_align.argtypes = [c_void_p, POINTER(c_int8), c_int32, c_uint8, c_uint8, c_uint8, c_uint16, c_int32, c_int32]
    lib.align_destroy.restype = None
    lib.align_destroy.argtypes = [POINTER(CAlignRes)]
    lib.cigar_int_to_len.restype = c_int32
    lib.cigar_int_to_len.argtypes = [c_int32]
    lib.cigar_int_to_op.restype = c_char
    lib.cigar_int_to_op.argtypes = [c_int32]
    lib.ssw_batch_create.restype = c_void_p
    lib.ssw_batch_create.argtypes = [c_int, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, POINTER(SSWScoring)]
    lib.ssw_batch_run.restype = c_int
    lib.ssw_batch_run.argtypes = [c_void_p]
    lib.ssw_batch_encode_ascii.restype = c_int
    lib.ssw_batch_encode_ascii.argtypes = [c_void_p]
    lib.ssw_batch_fetch.restype = c_int
    lib.ssw_batch_fetch.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_int64)]
    lib.ssw_batch_launch_count.restype = c_int64
    lib.ssw_batch_launch_count.argtypes = [c_void_p]
    lib.ssw_batch_stage_ms.restype = c_int
    lib.ssw_batch_stage_ms.argtypes = [c_void_p, c_void_p]
    lib.ssw_batch_destroy.restype = None
    lib.ssw_batch_destroy.argtypes = [c_void_p]
    lib.ssw_align_batch.restype = c_int
    lib.ssw_align_batch.argtypes = [c_int, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    POINTER(SSWScoring), c_void_p, c_void_p, c_int64, POINTER(c_int64)]
    lib.ssw_align_batch_multi.restype = c_int
    lib.ssw_align_batch_multi.argtypes = [c_void_p, c_int, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, POINTER(SSWScoring), c_void_p, c_void_p, c_int64, POINTER(c_int64)]
    lib.ssw_align_batch_multi_packed.restype = c_int
    lib.ssw_align_batch_multi_packed.argtypes = lib.ssw_align_batch_multi.argtypes
    lib.ssw_batch_create_packed.restype = c_void_p
    lib.ssw_batch_create_packed.argtypes = lib.ssw_batch_create.argtypes
    lib.ssw_pack_dna4.restype = None
    lib.ssw_pack_dna4.argtypes = [c_void_p, c_int64, c_void_p]
    lib.ssw_cuda_trim_pools.restype = c_int
    lib.ssw_cuda_last_error.restype = c_char_p
    lib.ssw_cuda_device_count.restype = c_int
    lib.ssw_cuda_dpx_peak.restype = c_int
    lib.ssw_cuda_dpx_peak.argtypes = [c_int, POINTER(c_double), POINTER(c_double)]
    return lib


def make_scoring(match, mismatch, gap_open, gap_extend, flag=1, filters=0, filterd=0):
    """5x5 matrix with a zero N row/column, as built by ssw_wrap.py:146-159."""
    sc = SSWScoring()
    m = [-mismatch] * 25
    for i in range(4):
        m[i * 5 + i] = match
    for i in range(5):
        m[4 * 5 + i] = 0
        m[i * 5 + 4] = 0
    sc.mat = (c_int8 * 25)(*m)
    sc.gap_open, sc.gap_extend, sc.flag = gap_open, gap_extend, flag
    sc.filters, sc.filterd = filters, filterd
    return sc


class SSWCudaError(RuntimeError):
    pass


def pack_dna4(codes):
    """codes 0..4 (one per byte) -> two bases per byte, low nibble first: the input of the *_packed entry points
    (half the host-to-device bytes; offsets and lengths keep counting bases)."""
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    out = np.zeros((len(codes) + 1) // 2, dtype=np.uint8)
    Aligner.libssw.ssw_pack_dna4(codes.ctypes.data, len(codes), out.ctypes.data)
    return out


# Opt-in: route this process's alignments through a GPU-owner service (server.py) instead of the local library.
# Meant for forked pool workers (find_bsj.py:338-345, collapse.py:842-851), which must not create CUDA contexts of
# their own: ``Aligner.align`` then blocks on the service, so unmodified call sites batch across workers.
_SERVICE = None


def use_service(client):
    """client: server.AlignClient (or None to go back to the in-process library)."""
    global _SERVICE
    _SERVICE = client


#~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~#
class DeviceBatch(object):
    """A batch of independent (query, reference) pairs resident in device memory (ssw_batch_*)."""

    libssw = _load()

    def __init__(self, seqs, q_off, q_len, r_off, r_len, match, mismatch, gap_open, gap_extend,
                 flag=1, mask_len=None, device=0, stream=None, filters=0, filterd=0, ascii=False, packed_bases=0):
        # packed_bases > 0: `seqs` is a 4-bit packed buffer (pack_dna4) holding that many bases
        self.seqs = np.ascontiguousarray(seqs, dtype=np.uint8 if packed_bases else np.int8)
        self.q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        self.q_len = np.ascontiguousarray(q_len, dtype=np.int32)
        self.r_off = np.ascontiguousarray(r_off, dtype=np.int64)
        self.r_len = np.ascontiguousarray(r_len, dtype=np.int32)
        self.mask_len = None if mask_len is None else np.ascontiguousarray(mask_len, dtype=np.int32)
        self.n = len(self.q_len)
        self.flag = flag
        self.scoring = make_scoring(match, mismatch, gap_open, gap_extend, flag, filters, filterd)
        create = self.libssw.ssw_batch_create_packed if packed_bases else self.libssw.ssw_batch_create
        self.handle = create(
            device, stream, self.n, self.seqs.ctypes.data, packed_bases if packed_bases else self.seqs.size,
            self.q_off.ctypes.data, self.q_len.ctypes.data, self.r_off.ctypes.data, self.r_len.ctypes.data,
            None if self.mask_len is None else self.mask_len.ctypes.data, byref(self.scoring))
        if not self.handle:
            raise SSWCudaError(self.libssw.ssw_cuda_last_error().decode())
        if ascii:
            # `seqs` holds the raw letters: the encode of ssw_wrap.py:234-252 runs on the device
            rc = self.libssw.ssw_batch_encode_ascii(self.handle)
            if rc != 0:
                raise SSWCudaError("ssw_batch_encode_ascii: %d %s" % (rc, self.libssw.ssw_cuda_last_error().decode()))

    @property
    def h2d_bytes(self):
        return int(self.seqs.nbytes + self.q_off.nbytes + self.q_len.nbytes + self.r_off.nbytes +
                   self.r_len.nbytes + 4 * self.n)

    def run(self):
        rc = self.libssw.ssw_batch_run(self.handle)
        if rc != 0:
            raise SSWCudaError("ssw_batch_run: %d %s" % (rc, self.libssw.ssw_cuda_last_error().decode()))

    def stage_ms(self):
        """device ms of (forward, deciding, reverse, cigar) stages of the last run()"""
        ms = np.zeros(4, dtype=np.float32)
        rc = self.libssw.ssw_batch_stage_ms(self.handle, ms.ctypes.data)
        if rc != 0:
            raise SSWCudaError("ssw_batch_stage_ms: %d" % rc)
        return ms

    def launch_count(self):
        return int(self.libssw.ssw_batch_launch_count(self.handle))

    def fetch(self, cigar_cap=None, out=None, cig=None):
        """-> (records[RESULT_DTYPE], cigar uint32 array).  `out` / `cig` may be preallocated (e.g. pinned)."""
        if out is None:
            out = np.zeros(self.n, dtype=RESULT_DTYPE)
        if cig is None:
            if cigar_cap is None:
                cigar_cap = int(2 * self.q_len.astype(np.int64).sum() + 3 * self.n + 16) if self.flag else 1
            cig = np.empty(cigar_cap, dtype=np.uint32)
        cigar_cap = len(cig)
        used = c_int64(0)
        rc = self.libssw.ssw_batch_fetch(self.handle, out.ctypes.data, cig.ctypes.data, cigar_cap, byref(used))
        if rc != 0:
            raise SSWCudaError("ssw_batch_fetch: %d %s" % (rc, self.libssw.ssw_cuda_last_error().decode()))
        self.d2h_bytes = int(out.nbytes + 4 * used.value)
        return out, cig[:used.value]

    def close(self):
        if getattr(self, "handle", None):
            self.libssw.ssw_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


#~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~#
class Aligner(object):
    """
    @class  SSWAligner
    @brief  Same interface as the reference wrapper (ssw_wrap.py:40-259), backed by libssw_cuda.so
    """

    # Dictionnary to map Nucleotide to int as expected by the SSW C library
    base_to_int = {'A': 0, 'C': 1, 'G': 2, 'T': 3, 'N': 4, 'a': 0, 'c': 1, 'g': 2, 't': 3, 'n': 4}
    int_to_base = {0: 'A', 1: 'C', 2: 'G', 3: 'T', 4: 'N'}

    # Load the library at class-body time, like the reference (import fails loudly if it is missing)
    libssw = _load()
    ssw_init = libssw.ssw_init
    init_destroy = libssw.init_destroy
    ssw_align = libssw.ssw_align
    align_destroy = libssw.align_destroy

    def __repr__(self):
        msg = self.__str__()
        msg += "SCORE PARAMETERS:\n"
        msg += " Gap Weight     Open: {}     Extension: {}\n".format(-self.gap_open, -self.gap_extend)
        msg += " Align Weight   Match: {}    Mismatch: {}\n\n".format(self.match, -self.mismatch)
        msg += "RESULT PARAMETERS:\n"
        msg += " Report cigar           {}\n".format(self.report_cigar)
        msg += " Report secondary match {}\n\n".format(self.report_secondary)
        msg += "REFERENCE SEQUENCE :\n"
        shown = min(self.ref_len, 50)
        msg += "".join([self.int_to_base[int(self.ref_seq[i])] for i in range(shown)])
        msg += "\n" if self.ref_len <= 50 else "...\n"
        msg += " Lenght :{} nucleotides\n".format(self.ref_len)
        return msg

    def __str__(self):
        return "\n<Instance of {} from {} >\n".format(self.__class__.__name__, self.__module__)

    def __init__(self, ref_seq="", match=2, mismatch=2, gap_open=3, gap_extend=1,
                 report_secondary=False, report_cigar=False):
        self.report_secondary = report_secondary
        self.report_cigar = report_cigar
        self.set_gap(gap_open, gap_extend)
        self.set_mat(match, mismatch)
        self.set_ref(ref_seq)

    #~~~~~~~SETTERS METHODS~~~~~~~#

    def set_gap(self, gap_open=3, gap_extend=1):
        self.gap_open = gap_open
        self.gap_extend = gap_extend

    def set_mat(self, match=2, mismatch=2):
        """5x5 matrix, ambiguous base: no penalty (ssw_wrap.py:146-159)."""
        self.match = match
        self.mismatch = mismatch
        self.mat = make_scoring(match, mismatch, 0, 0).mat

    def set_ref(self, ref_seq):
        self._ref_src = ref_seq
        if ref_seq is not None and len(ref_seq):
            self.ref_len = len(ref_seq)
            self.ref_seq = self._DNA_to_int_mat(ref_seq, self.ref_len)
        else:
            self.ref_len = 0
            self.ref_seq = ""

    #~~~~~~~PUBLIC METHODS~~~~~~~#

    def align(self, query_seq, min_score=0, min_len=0):
        """One pair through the legacy C ABI (ssw_init / ssw_align), same flow as ssw_wrap.py:174-230."""
        if _SERVICE is not None:
            return self._align_via_service(query_seq, min_score, min_len)
        query_len = len(query_seq)
        query_seq = self._DNA_to_int_mat(query_seq, query_len)
        profile = self.ssw_init(query_seq, c_int32(query_len), self.mat, 5, 2)
        mask_len = query_len // 2 if query_len > 30 else 15
        c_result = self.ssw_align(profile, self.ref_seq, c_int32(self.ref_len), self.gap_open, self.gap_extend,
                                  1, 0, 0, mask_len)
        if c_result and c_result.contents:
            score = c_result.contents.score
            match_len = c_result.contents.query_end - c_result.contents.query_begin + 1
        else:
            score = -999999999999
            match_len = -10000000000
        if score >= min_score and match_len >= min_len:
            py_result = PyAlignRes(c_result, query_len, self.report_secondary, self.report_cigar)
        else:
            py_result = None
        self._init_destroy(profile)
        if c_result:
            self._align_destroy(c_result)
        return py_result

    def align_batch(self, queries, min_score=0, min_len=0, device=0):
        """Many queries against this aligner's reference in one device batch.
        Returns a list with one PyAlignRes (or None when filtered / failed) per query, identical to
        ``[self.align(q, min_score, min_len) for q in queries]``."""
        ref = np.frombuffer(self.ref_seq, dtype=np.int8) if self.ref_len else np.zeros(0, np.int8)
        return align_pairs([ref] * len(queries), queries, self.match, self.mismatch, self.gap_open,
                           self.gap_extend, self.report_secondary, self.report_cigar, min_score, min_len,
                           device=device, _shared_ref=True)

    #~~~~~~~PRIVATE METHODS~~~~~~~#

    def _align_via_service(self, query_seq, min_score, min_len):
        """the same call, executed by the GPU-owner process (server.py); this process never touches the device"""
        as_str = lambda x: x if isinstance(x, str) else "".join("ACGTN"[min(int(c) & 0xff, 4)] for c in x)
        rec, cig = _SERVICE.align(as_str(self._ref_src), as_str(query_seq), self.match, self.mismatch, self.gap_open,
                                  self.gap_extend, need_cigar=True)
        if (rec['status'] & 0xff) not in (0, 1):
            return None
        if rec['score1'] >= min_score and rec['read_end1'] - rec['read_begin1'] + 1 >= min_len:
            return PyAlignRes.from_record(rec, cig, len(query_seq), self.report_secondary, self.report_cigar)
        return None

    def _DNA_to_int_mat(self, seq, len_seq):
        codes = encode_dna(seq)
        return (c_int8 * len_seq).from_buffer_copy(codes.tobytes())

    def _init_destroy(self, profile):
        self.init_destroy(profile)

    def _align_destroy(self, align):
        self.align_destroy(align)


#~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~#
class PyAlignRes(object):
    """
    @class  PyAlignRes
    @brief  Same attributes as the reference class (ssw_wrap.py:267-379)
    """

    libssw = Aligner.libssw
    cigar_int_to_len = libssw.cigar_int_to_len
    cigar_int_to_op = libssw.cigar_int_to_op

    def __repr__(self):
        msg = self.__str__()
        msg += "OPTIMAL MATCH\n"
        msg += "Score            {}\n".format(self.score)
        msg += "Reference begin  {}\n".format(self.ref_begin)
        msg += "Reference end    {}\n".format(self.ref_end)
        msg += "Query begin      {}\n".format(self.query_begin)
        msg += "Query end        {}\n".format(self.query_end)
        if self.cigar_string:
            msg += "Cigar_string     {}\n".format(self.cigar_string)
        if self.score2:
            msg += "SUB-OPTIMAL MATCH\n"
            msg += "Score 2           {}\n".format(self.score2)
            msg += "Ref_end2          {}\n".format(self.ref_end2)
        return msg

    def __str__(self):
        return "\n<Instance of {} from {} >\n".format(self.__class__.__name__, self.__module__)

    def __init__(self, Res, query_len, report_secondary=False, report_cigar=False):
        c = Res.contents
        self._fill(c.score, c.ref_begin, c.ref_end, c.query_begin, c.query_end, c.score2, c.ref_end2,
                   [c.cigar[i] for i in range(c.cigarLen)] if (report_cigar and c.cigarLen > 0) else None,
                   query_len, report_secondary, report_cigar)

    @classmethod
    def from_record(cls, rec, cigar, query_len, report_secondary=False, report_cigar=False):
        """Build from one row of DeviceBatch.fetch()."""
        self = cls.__new__(cls)
        ops = None
        if report_cigar and rec['cigar_len'] > 0:
            ops = cigar[rec['cigar_off']:rec['cigar_off'] + rec['cigar_len']]
        self._fill(int(rec['score1']), int(rec['ref_begin1']), int(rec['ref_end1']), int(rec['read_begin1']),
                   int(rec['read_end1']), int(rec['score2']), int(rec['ref_end2']), ops, query_len,
                   report_secondary, report_cigar)
        return self

    def _fill(self, score, ref_begin, ref_end, query_begin, query_end, score2, ref_end2, ops, query_len,
              report_secondary, report_cigar):
        self.score = score
        self.ref_begin = ref_begin
        self.ref_end = ref_end
        self.query_begin = query_begin
        self.query_end = query_end
        if report_secondary and score2 != 0:
            self.score2 = score2
            self.ref_end2 = ref_end2
        else:
            self.score2 = None
            self.ref_end2 = None
        if report_cigar and ops is not None and len(ops) > 0:
            self.cigar_string = self._cigar_string(ops, len(ops), query_len)
        else:
            self.cigar_string = None

    def _cigar_string(self, cigar, cigar_len, query_len):
        """SAM-like string with soft clips, ssw_wrap.py:349-379 (ops decoded in Python, not per-op ctypes calls)."""
        parts = []
        if self.query_begin > 0:
            parts.append('{}S'.format(self.query_begin))
        for i in range(cigar_len):
            c = int(cigar[i])
            code = c & 0xF
            parts.append('{}{}'.format(c >> 4, "MIDNSHP=X"[code] if code < 9 else 'M'))
        end_len = query_len - self.query_end - 1
        if end_len != 0:
            parts.append('{}S'.format(end_len))
        return "".join(parts)


#~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~#
def _encode_many(seqs):
    """Encode a list of sequences with ONE LUT pass over their concatenation; returns (codes, lengths)."""
    if all(isinstance(x, str) for x in seqs):
        lens = np.fromiter(map(len, seqs), dtype=np.int64, count=len(seqs))
        blob = "".join(seqs).encode("latin-1", "replace").translate(_ENC_TABLE)     # C-speed table lookup
        return np.frombuffer(blob, dtype=np.int8), lens
    parts = [encode_dna(x) for x in seqs]
    lens = np.fromiter(map(len, parts), dtype=np.int64, count=len(parts))
    return (np.concatenate(parts) if parts else np.zeros(0, np.int8)), lens


def pack_pairs(refs, queries, shared_ref=False, raw=False):
    """Concatenate the sequences into the struct-of-arrays layout of ssw_batch_create (references first, then
    queries; with ``shared_ref`` the one reference is stored once).  Returns (seqs, q_off, q_len, r_off, r_len);
    with ``raw=True`` a sixth item says whether ``seqs`` still holds ASCII letters (all inputs were ``str``: one
    join, no host-side table pass -- the device encodes them, DeviceBatch(ascii=True))."""
    refs = list(refs[:1]) if shared_ref else list(refs)
    queries = list(queries)
    is_ascii = False
    if all(isinstance(x, str) for x in refs) and all(isinstance(x, str) for x in queries):
        r_lens = np.fromiter(map(len, refs), dtype=np.int64, count=len(refs))
        q_lens = np.fromiter(map(len, queries), dtype=np.int64, count=len(queries))
        blob = "".join(refs + queries).encode("latin-1", "replace")
        if raw:
            is_ascii = True
        else:
            blob = blob.translate(_ENC_TABLE)
        seqs = np.frombuffer(blob, dtype=np.int8)
    else:
        r_codes, r_lens = _encode_many(refs)
        q_codes, q_lens = _encode_many(queries)
        seqs = np.concatenate([r_codes, q_codes]) if (len(r_codes) or len(q_codes)) else np.zeros(0, np.int8)
    nq = len(q_lens)
    if shared_ref:
        r_off = np.zeros(nq, dtype=np.int64)
        r_len = np.full(nq, int(r_lens[0]) if len(r_lens) else 0, dtype=np.int32)
    else:
        r_off = (np.cumsum(r_lens) - r_lens).astype(np.int64)
        r_len = r_lens.astype(np.int32)
    q_off = (int(r_lens.sum()) + np.cumsum(q_lens) - q_lens).astype(np.int64)
    out = (np.ascontiguousarray(seqs, dtype=np.int8), q_off, q_lens.astype(np.int32), r_off, r_len)
    return out + (is_ascii,) if raw else out


def align_pairs(refs, queries, match=2, mismatch=2, gap_open=3, gap_extend=1, report_secondary=False,
                report_cigar=False, min_score=0, min_len=0, device=0, need_cigar=None, _shared_ref=False,
                as_records=False):
    """Batched equivalent of ``[Aligner(r, ...).align(q, min_score, min_len) for r, q in zip(refs, queries)]``.

    ``need_cigar`` defaults to ``report_cigar``: callers that never read ``cigar_string`` (every CIRI-long
    site except collapse.py:373-387) may skip the CIGAR pass; begin coordinates are still computed.
    ``as_records=True`` returns ``(records, cigar_ops)`` -- the RESULT_DTYPE array of the C interface and the
    uint32 op buffer its ``cigar_off`` / ``cigar_len`` index -- instead of one Python object per pair, which
    is what bounds this call for large batches (min_score / min_len are then left to the caller)."""
    if len(refs) != len(queries):
        raise ValueError("refs and queries differ in length")
    if not len(queries):
        return (np.zeros(0, dtype=RESULT_DTYPE), np.zeros(0, dtype=np.uint32)) if as_records else []
    # str inputs: upload the raw letters and let the device encode them (ssw_batch_encode_ascii);
    # SSW_CUDA_DEVICE_ENCODE=0 keeps the table pass on the host
    if os.environ.get("SSW_CUDA_DEVICE_ENCODE", "1") != "0":
        seqs, q_off, q_len, r_off, r_len, is_ascii = pack_pairs(refs, queries, _shared_ref, raw=True)
    else:
        seqs, q_off, q_len, r_off, r_len = pack_pairs(refs, queries, _shared_ref)
        is_ascii = False
    if need_cigar is None:
        need_cigar = report_cigar
    if _SERVICE is not None and all(isinstance(x, str) for x in refs) and all(isinstance(x, str) for x in queries):
        rec, cig = _SERVICE.align_pairs(refs, queries, match, mismatch, gap_open, gap_extend, need_cigar=need_cigar)
        if as_records:
            return rec, cig
        q_len = [len(q) for q in queries]
        return [None if (r['status'] & 0xff) not in (0, 1) or r['score1'] < min_score or r['read_end1'] - r['read_begin1'] + 1 < min_len
                else PyAlignRes.from_record(r, cig, q_len[i], report_secondary, report_cigar) for i, r in enumerate(rec)]
    # flag 1 = begin + CIGAR (what the reference wrapper always asks for); flag 8 is not defined by the
    # reference -- the begin-only mode is expressed with the distance filter: bit 2 with filterd < 0
    flag = 1 if need_cigar else 4
    with DeviceBatch(seqs, q_off, q_len, r_off, r_len, match, mismatch, gap_open, gap_extend, flag=flag,
                     device=device, filterd=0 if need_cigar else -1, ascii=is_ascii) as b:
        b.run()
        rec, cig = b.fetch(cigar_cap=None if need_cigar else 1)     # coordinate-only mode: no CIGAR buffer on either side
    if as_records:
        return rec, cig
    out = []
    for i in range(len(queries)):
        r = rec[i]
        # status 1 = the traceback left the band (ssw.c:642-673 reads unwritten direction bytes there): score and
        # coordinates are exact, the CIGAR is empty -> cigar_string None
        if (r['status'] & 0xff) not in (0, 1):
            out.append(None)
            continue
        match_len = r['read_end1'] - r['read_begin1'] + 1
        if r['score1'] >= min_score and match_len >= min_len:
            out.append(PyAlignRes.from_record(r, cig, int(q_len[i]), report_secondary, report_cigar))
        else:
            out.append(None)
    return out


def align_arrays(seqs, q_off, q_len, r_off, r_len, match, mismatch, gap_open, gap_extend, flag=1, mask_len=None,
                 device=0, out=None, cig=None, devices=None, filters=0, filterd=0, packed_bases=0):
    """The one-shot C call (ssw_align_batch / ssw_align_batch_multi) on struct-of-arrays host buffers: upload, all
    kernels and download in one call, chunk-pipelined inside the library.  ``devices=[0, 1, ...]`` spreads the
    chunks over several GPUs of the box (one host thread per device, results gathered at the pairs' indices;
    bit-identical to the single-device call).  ``packed_bases=n``: `seqs` is a 4-bit packed buffer (pack_dna4) of n
    bases -- half the upload.  Returns (records, cigar ops)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8 if packed_bases else np.int8)
    q_off = np.ascontiguousarray(q_off, dtype=np.int64); r_off = np.ascontiguousarray(r_off, dtype=np.int64)
    q_len = np.ascontiguousarray(q_len, dtype=np.int32); r_len = np.ascontiguousarray(r_len, dtype=np.int32)
    n = len(q_len)
    if out is None:
        out = np.zeros(n, dtype=RESULT_DTYPE)
    if cig is None:
        cig = np.empty(int(2 * q_len.astype(np.int64).sum() + 3 * n + 16) if flag else 1, dtype=np.uint32)
    sc = make_scoring(match, mismatch, gap_open, gap_extend, flag, filters, filterd)
    used = c_int64(0)
    ml = None if mask_len is None else np.ascontiguousarray(mask_len, dtype=np.int32)
    devs = np.ascontiguousarray([device] if devices is None else list(devices), dtype=np.int32)
    call = Aligner.libssw.ssw_align_batch_multi_packed if packed_bases else Aligner.libssw.ssw_align_batch_multi
    rc = call(devs.ctypes.data, len(devs), n, seqs.ctypes.data, packed_bases if packed_bases else seqs.size, q_off.ctypes.data,
                                              q_len.ctypes.data, r_off.ctypes.data, r_len.ctypes.data,
                                              None if ml is None else ml.ctypes.data, byref(sc), out.ctypes.data,
                                              cig.ctypes.data, len(cig), byref(used))
    if rc != 0:
        raise SSWCudaError("ssw_align_batch_multi: %d %s" % (rc, Aligner.libssw.ssw_cuda_last_error().decode()))
    return out, cig[:used.value]


def dpx_peak(device=0):
    """(lane-instructions/s of VIADDMNMX.S16x2, SM clock MHz) measured on `device`."""
    v, mhz = c_double(0), c_double(0)
    rc = Aligner.libssw.ssw_cuda_dpx_peak(device, byref(v), byref(mhz))
    if rc != 0:
        raise SSWCudaError(Aligner.libssw.ssw_cuda_last_error().decode())
    return v.value, mhz.value
