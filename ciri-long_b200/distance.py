"""Batched edit distance on the GPU: the drop-in for ``CIRI_long.utils.distance`` (utils.py:153-159).

The reference computes ``Levenshtein.distance(x, y)`` when either string has at most 50 symbols and
``edlib.align(x, y)['editDistance']`` (global alignment, unit costs) otherwise -- the same number.  It is
called once per pair from ``collapse.avg_score`` (collapse.py:156-158) and from the O(k^2) loop of
``collapse.cluster_sequence`` (collapse.py:466-473).  Here the pairs of a whole loop go to the device in one
``ssw_cuda_edit_distance_batch`` call (bit-parallel Myers kernels, ciri-long_b200/csrc/edit_distance.cu).
There is no CPU implementation behind these functions.
"""
from ctypes import c_int, c_int32, c_int64, c_void_p

import numpy as np

from .ssw_wrap import Aligner, SSWCudaError

_lib = Aligner.libssw
_lib.ssw_cuda_edit_distance_batch.restype = c_int
_lib.ssw_cuda_edit_distance_batch.argtypes = [c_int, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p]


def distance_arrays(seqs, x_off, x_len, y_off, y_len, device=0):
    """Struct-of-arrays form: ``seqs`` uint8 bytes, pair p = seqs[x_off[p]:+x_len[p]] vs seqs[y_off[p]:+y_len[p]]."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    x_off = np.ascontiguousarray(x_off, dtype=np.int64); y_off = np.ascontiguousarray(y_off, dtype=np.int64)
    x_len = np.ascontiguousarray(x_len, dtype=np.int32); y_len = np.ascontiguousarray(y_len, dtype=np.int32)
    n = len(x_len)
    out = np.zeros(n, dtype=np.int32)
    rc = _lib.ssw_cuda_edit_distance_batch(device, n, seqs.ctypes.data, seqs.size, x_off.ctypes.data, x_len.ctypes.data,
                                           y_off.ctypes.data, y_len.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise SSWCudaError("ssw_cuda_edit_distance_batch: %d %s" % (rc, _lib.ssw_cuda_last_error().decode()))
    return out


def _as_bytes(s):
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


def distance_batch(xs, ys, device=0):
    """``[distance(x, y) for x, y in zip(xs, ys)]`` in one device call."""
    if len(xs) != len(ys):
        raise ValueError("xs and ys differ in length")
    if not len(xs):
        return np.zeros(0, dtype=np.int32)
    # every distinct string is stored once (the cluster loops compare each sequence with many others)
    blobs, where = [], {}
    pos = 0

    def place(s):
        nonlocal pos
        b = _as_bytes(s)
        hit = where.get(b)
        if hit is None:
            hit = where[b] = pos
            blobs.append(b)
            pos += len(b)
        return hit, len(b)

    xo, xl, yo, yl = [], [], [], []
    for x, y in zip(xs, ys):
        o, l = place(x); xo.append(o); xl.append(l)
        o, l = place(y); yo.append(o); yl.append(l)
    seqs = np.frombuffer(b"".join(blobs) or b"\0", dtype=np.uint8)
    return distance_arrays(seqs, xo, xl, yo, yl, device)


def distance(x, y, device=0):
    """Drop-in for ``utils.distance(x, y)`` (one pair per call: compatibility, not throughput)."""
    return int(distance_batch([x], [y], device)[0])


def cluster_distance_matrix(sequences, device=0):
    """The distance matrix of ``collapse.cluster_sequence`` (collapse.py:466-473):
    ``dist[i][j] = distance(s_i, s_j) / max(len(s_i), len(s_j))`` for i <= j, then ``dist + dist.T``."""
    k = len(sequences)
    ii, jj = np.triu_indices(k)
    d = distance_batch([sequences[i] for i in ii], [sequences[j] for j in jj], device).astype(np.float64)
    lens = np.array([len(s) for s in sequences], dtype=np.float64)
    dist = np.zeros((k, k))
    dist[ii, jj] = d / np.maximum(np.maximum(lens[ii], lens[jj]), 1)
    return dist + dist.T
