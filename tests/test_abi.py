"""CPU tests of the boundary: libssw_cuda.so loads without a GPU, exports every symbol the header
declares, keeps the reference struct layout, and fails loudly (never silently on the CPU) when no
device is present.  No alignment is computed here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sw():
    subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "ciri-long_b200", "csrc")])
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import ssw_wrap
    return ssw_wrap


def test_exports_every_declared_symbol(sw):
    hdr = open(os.path.join(ROOT, "include", "ssw_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(ssw_\w+|init_destroy|align_destroy|cigar_int_to_op|cigar_int_to_len)\s*\(", hdr))
    names -= {"ssw_scoring", "ssw_result", "ssw_batch"}
    assert {"ssw_init", "ssw_align", "init_destroy", "align_destroy", "cigar_int_to_op", "cigar_int_to_len",
            "ssw_batch_create", "ssw_batch_run", "ssw_batch_fetch", "ssw_batch_destroy", "ssw_align_batch", "ssw_align_batch_multi",
            "ssw_batch_stage_ms", "ssw_encode_dna", "ssw_cuda_dpx_peak", "ssw_batch_create_packed", "ssw_align_batch_multi_packed",
            "ssw_pack_dna4", "ssw_cuda_trim_pools"} <= names
    lib = ctypes.CDLL(os.path.join(ROOT, "ciri-long_b200", "libssw_cuda.so"))
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_struct_layouts(sw):
    # s_align of the reference: 40 bytes, offsets 0,2,4,8,12,16,20,24,32 (ssw.h:42-52)
    C = sw.CAlignRes
    assert ctypes.sizeof(C) == 40
    assert [getattr(C, f).offset for f, _ in C._fields_] == [0, 2, 4, 8, 12, 16, 20, 24, 32]
    assert ctypes.sizeof(sw.SSWResult) == sw.RESULT_DTYPE.itemsize == 48
    for f, _ in sw.SSWResult._fields_:
        assert getattr(sw.SSWResult, f).offset == sw.RESULT_DTYPE.fields[f][1]
    assert ctypes.sizeof(sw.SSWScoring) == 40 and sw.SSWScoring.filters.offset == 30 and sw.SSWScoring.filterd.offset == 36


def test_encode_and_cigar_helpers(sw):
    lib = sw.Aligner.libssw
    s = b"ACGTNacgtnXR-*"
    out = np.zeros(len(s), dtype=np.int8)
    lib.ssw_encode_dna.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_void_p]
    lib.ssw_encode_dna(s, len(s), out.ctypes.data)
    assert out.tolist() == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 4, 4, 4, 4]
    assert sw.encode_dna(s.decode()).tolist() == out.tolist()
    for code, ch in enumerate("MIDNSHP=X"):
        assert lib.cigar_int_to_op((37 << 4) | code) == ch.encode()
        assert lib.cigar_int_to_len((37 << 4) | code) == 37
    assert lib.cigar_int_to_op((5 << 4) | 12) == b"M"          # unknown codes map to M (ssw.c:891-893)


def test_no_device_fails_loudly(sw, capfd):
    """On the CPU-only build box every alignment entry point must refuse, not fall back."""
    if sw.Aligner.libssw.ssw_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    al = sw.Aligner("ACGTACGTACGT", 1, 1, 1, 1)
    assert al.align("ACGT") is None                           # NULL from ssw_align -> None, like the reference
    assert "no usable CUDA device" in capfd.readouterr().err
    with pytest.raises(sw.SSWCudaError):
        sw.align_pairs(["ACGTACGT"], ["ACGT"], 1, 1, 1, 1)


def test_unsupported_scoring_is_reported(sw):
    sc = sw.make_scoring(1, 3, 1, 1)                           # 2*gap_extend < mismatch: outside the exact kernels
    h = sw.Aligner.libssw.ssw_batch_create(0, None, 0, None, 0, None, None, None, None, None, ctypes.byref(sc))
    assert not h
    assert b"not supported" in sw.Aligner.libssw.ssw_cuda_last_error()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "ciri-long_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_pack_pairs_and_scoring(sw):
    seqs, q_off, q_len, r_off, r_len = sw.pack_pairs(["ACGT", "TTNN"], ["AC", "GTACG"])
    assert q_len.tolist() == [2, 5] and r_len.tolist() == [4, 4]
    assert seqs[q_off[1]:q_off[1] + 5].tolist() == [2, 3, 0, 1, 2] and seqs[r_off[1]:r_off[1] + 4].tolist() == [3, 3, 4, 4]
    seqs, q_off, q_len, r_off, r_len = sw.pack_pairs(["ACGT"] * 3, ["A", "CC", "GGG"], shared_ref=True)
    assert r_off.tolist() == [0, 0, 0] and r_len.tolist() == [4, 4, 4] and q_off.tolist() == [4, 5, 7]
    m = list(sw.make_scoring(10, 4, 8, 2).mat)
    assert m[0] == 10 and m[1] == -4 and m[4] == 0 and m[20:25] == [0] * 5


def test_revcomp_matches_the_reference_table(sw):
    """CIRI_long/utils.py:118-120 complements upper-case A/T/C/G only: lower-case (soft-masked genome) bases are
    reversed but NOT complemented, N stays N.  The minus-strand windows of align_clip_segments_batch
    (find_bsj.py:214) must be the same string the reference aligns against."""
    from ciri_long_b200 import callsites as cs
    ref_table = str.maketrans("ATCG", "TAGC")                   # the reference's own table, restated
    for seq in ("ACGTN", "acgtn", "AAccGGttNn", "GATTACAgattacaNNNN", ""):
        assert cs.revcomp(seq) == seq.translate(ref_table)[::-1]
    assert cs.revcomp("AAccGG") == "CCccTT"                     # lower case reversed, not complemented
    window = "ACGTacgtNNACGT"
    assert cs.revcomp(window) == "ACGTNNtgcaACGT"


def test_pack_dna4_layout(sw):
    """two bases per byte, low nibble first; codes above 4 become N (4); odd lengths leave the last high nibble 0"""
    c = np.array([0, 1, 2, 3, 4, 3, 2], dtype=np.int8)
    assert sw.pack_dna4(c).tolist() == [0 | 1 << 4, 2 | 3 << 4, 4 | 3 << 4, 2]
    assert sw.pack_dna4(np.array([7, -1], dtype=np.int8)).tolist() == [4 | 4 << 4]
