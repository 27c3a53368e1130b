"""Synthetic SSW workloads of the shapes named in BASELINE.json / SURVEY.md section 8(d).

Everything is generated as struct-of-arrays over one concatenated int8 code buffer
(A C G T N -> 0..4, the encoding of ssw_wrap.py:50), which is the layout the batched C ABI
(`ssw_align_batch`, include/ssw_cuda.h) consumes.  All generators are vectorised over the whole batch
so that a million pairs take seconds, and are seeded (`numpy.random.default_rng(20261017 + config_id)`).
"""
from dataclasses import dataclass

import numpy as np

SEED_BASE = 20261017


@dataclass
class PairBatch:
    """n pairs over one code buffer.  query = "read" of ssw_align, ref = its "ref"."""
    seqs: np.ndarray      # int8 codes, concatenated
    q_off: np.ndarray     # int64
    q_len: np.ndarray     # int32
    r_off: np.ndarray     # int64
    r_len: np.ndarray     # int32
    match: int
    mismatch: int
    gap_open: int
    gap_extend: int
    name: str = ""

    def __len__(self):
        return len(self.q_len)

    @property
    def cells(self):
        return int((self.q_len.astype(np.int64) * self.r_len.astype(np.int64)).sum())

    def query(self, i):
        return self.seqs[self.q_off[i]:self.q_off[i] + self.q_len[i]]

    def ref(self, i):
        return self.seqs[self.r_off[i]:self.r_off[i] + self.r_len[i]]

    def subset(self, idx):
        """Re-pack a subset of pairs into a fresh buffer."""
        idx = np.asarray(idx)
        ql, rl = self.q_len[idx], self.r_len[idx]
        tot = np.concatenate([[0], np.cumsum(np.stack([ql, rl], 1).reshape(-1).astype(np.int64))])
        seqs = np.empty(tot[-1], dtype=np.int8)
        q_off, r_off = tot[0:-1:2].copy(), tot[1::2].copy()
        for k, i in enumerate(idx):
            seqs[q_off[k]:q_off[k] + ql[k]] = self.query(i)
            seqs[r_off[k]:r_off[k] + rl[k]] = self.ref(i)
        return PairBatch(seqs, q_off, ql.copy(), r_off, rl.copy(), self.match, self.mismatch,
                         self.gap_open, self.gap_extend, self.name + "[subset]")


def _ragged_arange(starts, lens):
    """concatenate [arange(s, s+l) for s, l in zip(starts, lens)] without a Python loop."""
    lens = lens.astype(np.int64)
    total = int(lens.sum())
    seg_start = np.cumsum(lens) - lens
    return np.repeat(starts.astype(np.int64) - seg_start, lens) + np.arange(total, dtype=np.int64)


def noisy_channel(codes, seg_len, rng, sub=0.05, ins=0.04, dele=0.04, max_run=3, n_frac=0.0):
    """ONT-like channel applied to concatenated segments: substitutions, deletions, insertion runs
    U[1, max_run] after a base, then a fraction of bases replaced by N.  Returns (codes, seg_len)."""
    n = len(codes)
    seg_id = np.repeat(np.arange(len(seg_len)), seg_len)
    u = rng.random(n)
    keep = u >= dele
    is_sub = (u >= dele) & (u < dele + sub)
    out = codes.copy()
    out[is_sub] = (out[is_sub] + rng.integers(1, 4, size=int(is_sub.sum()))) % 4
    runs = np.where(rng.random(n) < ins, rng.integers(1, max_run + 1, size=n), 0)
    reps = keep.astype(np.int64) + runs
    new = np.repeat(out, reps)
    # positions that are inserted bases: within each repeated group, every copy after the kept one
    grp_start = np.cumsum(reps) - reps
    pos_in_grp = np.arange(len(new)) - np.repeat(grp_start, reps)
    inserted = pos_in_grp >= np.repeat(keep.astype(np.int64), reps)
    new[inserted] = rng.integers(0, 4, size=int(inserted.sum()))
    new_seg = np.repeat(seg_id, reps)
    if n_frac > 0:
        new[rng.random(len(new)) < n_frac] = 4
    new_len = np.bincount(new_seg, minlength=len(seg_len)).astype(np.int32)
    return new.astype(np.int8), new_len


def _pack(queries, q_len, refs, r_len):
    """Interleave per-pair query and ref blocks in one buffer: [q0 r0 q1 r1 ...]."""
    n = len(q_len)
    lens = np.stack([q_len.astype(np.int64), r_len.astype(np.int64)], 1).reshape(-1)
    offs = np.concatenate([[0], np.cumsum(lens)])
    q_off, r_off = offs[0:-1:2].copy(), offs[1::2].copy()
    seqs = np.empty(offs[-1], dtype=np.int8)
    seqs[_ragged_arange(q_off, q_len)] = queries
    seqs[_ragged_arange(r_off, r_len)] = refs
    return seqs, q_off, r_off


def bsj_refinement_pairs(n_pairs, seed=SEED_BASE + 2, ref_len=2000, q_min=300, q_max=800,
                         n_frac=0.01, params=(1, 1, 1, 1)):
    """Config C2: consensus segment (300-800 nt, ONT-like noise, 1 % N) vs a 2 kb genomic flank,
    find_bsj scoring 1/1/1/1 (find_bsj.py:204)."""
    rng = np.random.default_rng(seed)
    refs = rng.integers(0, 4, size=(n_pairs, ref_len), dtype=np.int8)
    ql0 = rng.integers(q_min, q_max + 1, size=n_pairs)
    start = (rng.random(n_pairs) * (ref_len - ql0 + 1)).astype(np.int64)
    src = _ragged_arange(start + np.arange(n_pairs, dtype=np.int64) * ref_len, ql0)
    q_codes, q_len = noisy_channel(refs.reshape(-1)[src], ql0, rng, n_frac=n_frac)
    r_len = np.full(n_pairs, ref_len, dtype=np.int32)
    seqs, q_off, r_off = _pack(q_codes, q_len, refs.reshape(-1), r_len)
    return PairBatch(seqs, q_off, q_len, r_off, r_len, *params, name="C2-bsj-refinement")


def rolling_circle_pairs(n_reads, seed=SEED_BASE + 3, read_min=2000, read_max=6000, copies_min=2,
                         copies_max=8, params=(10, 4, 8, 2)):
    """Config C3: NanoSim-style rolling-circle reads (misc/NanoSim.ipynb cells 2-3); pairs are
    (segment 0 as ref, segment k as query), collapse.py:259 scoring 10/4/8/2."""
    rng = np.random.default_rng(seed)
    read_len = rng.integers(read_min, read_max + 1, size=n_reads)
    copies = rng.integers(copies_min, copies_max + 1, size=n_reads)
    unit_len = np.maximum(read_len // copies, 20)
    n_seg = copies
    seg_unit = np.repeat(np.arange(n_reads), n_seg)            # which read each segment belongs to
    seg_len0 = unit_len[seg_unit]
    unit_off = np.cumsum(unit_len) - unit_len
    units = rng.integers(0, 4, size=int(unit_len.sum()), dtype=np.int8)
    rot = (rng.random(n_reads) * unit_len).astype(np.int64)
    # every segment is the unit rotated by the read's offset
    idx_in = _ragged_arange(np.zeros(len(seg_len0), dtype=np.int64), seg_len0)
    seg_rep = np.repeat(np.arange(len(seg_len0)), seg_len0)
    src = unit_off[seg_unit][seg_rep] + (idx_in + rot[seg_unit][seg_rep]) % unit_len[seg_unit][seg_rep]
    seg_codes, seg_len = noisy_channel(units[src], seg_len0, rng)
    seg_off = np.cumsum(seg_len.astype(np.int64)) - seg_len
    first_seg = np.cumsum(n_seg) - n_seg
    is_query = np.ones(len(seg_len), dtype=bool)
    is_query[first_seg] = False
    q_idx = np.nonzero(is_query)[0]
    r_idx = first_seg[seg_unit[q_idx]]
    ok = (seg_len[q_idx] > 0) & (seg_len[r_idx] > 0)
    q_idx, r_idx = q_idx[ok], r_idx[ok]
    return PairBatch(seg_codes, seg_off[q_idx], seg_len[q_idx].astype(np.int32), seg_off[r_idx],
                     seg_len[r_idx].astype(np.int32), *params, name="C3-rolling-circle")


def square_pairs(n_pairs, length, seed=SEED_BASE + 4, params=(1, 1, 1, 1), noise=True):
    """Config C4: m ~ n ~ length (length sweep 64..4096)."""
    rng = np.random.default_rng(seed + length)
    refs = rng.integers(0, 4, size=(n_pairs, length), dtype=np.int8)
    ql0 = np.full(n_pairs, length, dtype=np.int64)
    if noise:
        q_codes, q_len = noisy_channel(refs.reshape(-1), ql0, rng)
    else:
        q_codes, q_len = refs.reshape(-1).copy(), ql0.astype(np.int32)
    r_len = np.full(n_pairs, length, dtype=np.int32)
    seqs, q_off, r_off = _pack(q_codes, q_len, refs.reshape(-1), r_len)
    return PairBatch(seqs, q_off, q_len, r_off, r_len, *params, name="C4-square-%d" % length)


def junction_pairs(n_pairs, seed=SEED_BASE + 5, q_min=40, q_max=60, ref_len=20, params=(10, 4, 8, 2)):
    """S2-like tiny pairs (collapse.py:165-172): ~50-nt junction consensus vs a 20-nt genomic junction."""
    rng = np.random.default_rng(seed)
    ql = rng.integers(q_min, q_max + 1, size=n_pairs)
    q = rng.integers(0, 4, size=int(ql.sum()), dtype=np.int8)
    q_off = np.cumsum(ql) - ql
    st = (rng.random(n_pairs) * (ql - ref_len + 1)).astype(np.int64)
    src = _ragged_arange(q_off + st, np.full(n_pairs, ref_len))
    r_codes, r_len = noisy_channel(q[src], np.full(n_pairs, ref_len, dtype=np.int64), rng)
    seqs, q_off2, r_off = _pack(q, ql.astype(np.int32), r_codes, r_len)
    return PairBatch(seqs, q_off2, ql.astype(np.int32), r_off, r_len, *params, name="S2-junction")


def overflow_boundary_pairs(params=(1, 1, 1, 1), lengths=range(246, 262), seed=SEED_BASE + 6, flank=40):
    """Perfect matches of length L embedded in random flanks so that max+bias lands on 254/255/256
    (SURVEY 8c(iii)): the int8 -> int16 re-run boundary."""
    rng = np.random.default_rng(seed)
    qs, rs = [], []
    for L in lengths:
        core = rng.integers(0, 4, size=L, dtype=np.int8)
        # flanks chosen from a disjoint alphabet pattern so they do not extend the match
        rs.append(np.concatenate([(core[:1] + 1) % 4 * np.ones(flank, np.int8), core,
                                  (core[-1:] + 1) % 4 * np.ones(flank, np.int8)]).astype(np.int8))
        qs.append(np.concatenate([(core[:1] + 2) % 4 * np.ones(7, np.int8), core,
                                  (core[-1:] + 2) % 4 * np.ones(5, np.int8)]).astype(np.int8))
    q_len = np.array([len(x) for x in qs], dtype=np.int32)
    r_len = np.array([len(x) for x in rs], dtype=np.int32)
    seqs, q_off, r_off = _pack(np.concatenate(qs), q_len, np.concatenate(rs), r_len)
    return PairBatch(seqs, q_off, q_len, r_off, r_len, *params, name="overflow-boundary")


def from_lists(queries, refs, params, name="list"):
    """Build a batch from Python lists of int8 code arrays."""
    q_len = np.array([len(x) for x in queries], dtype=np.int32)
    r_len = np.array([len(x) for x in refs], dtype=np.int32)
    seqs, q_off, r_off = _pack(np.concatenate(queries).astype(np.int8), q_len,
                               np.concatenate(refs).astype(np.int8), r_len)
    return PairBatch(seqs, q_off, q_len, r_off, r_len, *params, name=name)


def bsj_refinement_pairs_torch(n_pairs, device, seed=SEED_BASE + 2, ref_len=2000, q_min=300, q_max=800,
                               n_frac=0.01, params=(1, 1, 1, 1), chunk=131072):
    """Config C2 generated on the GPU with torch (same recipe as bsj_refinement_pairs, different RNG
    stream): a million pairs take about a second instead of minutes.  Returns a PairBatch of numpy
    arrays (host)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    parts_q, parts_r, qlens = [], [], []
    for c0 in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - c0)
        refs = torch.randint(0, 4, (n, ref_len), dtype=torch.int8, device=device, generator=g)
        ql0 = torch.randint(q_min, q_max + 1, (n,), device=device, generator=g)
        start = (torch.rand(n, device=device, generator=g) * (ref_len - ql0 + 1).float()).long()
        seg_start = torch.cumsum(ql0, 0) - ql0
        total = int(ql0.sum())
        base = torch.repeat_interleave(start + torch.arange(n, device=device) * ref_len - seg_start, ql0)
        src = base + torch.arange(total, device=device)
        codes = refs.reshape(-1)[src]
        seg_id = torch.repeat_interleave(torch.arange(n, device=device), ql0)
        u = torch.rand(total, device=device, generator=g)
        keep = u >= 0.04
        is_sub = keep & (u < 0.09)
        shift = torch.randint(1, 4, (total,), dtype=torch.int8, device=device, generator=g)
        codes = torch.where(is_sub, (codes + shift) % 4, codes)
        runs = torch.where(torch.rand(total, device=device, generator=g) < 0.04,
                           torch.randint(1, 4, (total,), device=device, generator=g), torch.zeros((), dtype=torch.long, device=device))
        reps = keep.long() + runs
        new = torch.repeat_interleave(codes, reps)
        grp_start = torch.cumsum(reps, 0) - reps
        pos = torch.arange(new.numel(), device=device) - torch.repeat_interleave(grp_start, reps)
        inserted = pos >= torch.repeat_interleave(keep.long(), reps)
        rnd = torch.randint(0, 4, (new.numel(),), dtype=torch.int8, device=device, generator=g)
        new = torch.where(inserted, rnd, new)
        if n_frac > 0:
            new = torch.where(torch.rand(new.numel(), device=device, generator=g) < n_frac,
                              torch.full((), 4, dtype=torch.int8, device=device), new)
        new_len = torch.bincount(torch.repeat_interleave(seg_id, reps), minlength=n)
        parts_q.append(new.cpu().numpy())
        parts_r.append(refs.reshape(-1).cpu().numpy())
        qlens.append(new_len.cpu().numpy().astype(np.int32))
        del refs, codes, new, u, reps, pos, inserted, rnd, src, base, seg_id
    q_len = np.concatenate(qlens)
    r_len = np.full(n_pairs, ref_len, dtype=np.int32)
    # layout: per generation chunk [queries of the chunk][references of the chunk], so that consecutive
    # pairs reference a compact byte range (the one-shot C call uploads chunk by chunk)
    blocks, q_off, r_off, pos, p0 = [], np.empty(n_pairs, np.int64), np.empty(n_pairs, np.int64), 0, 0
    for qc, rc, ql in zip(parts_q, parts_r, qlens):
        k = len(ql)
        q_off[p0:p0 + k] = pos + np.cumsum(ql.astype(np.int64)) - ql
        pos += len(qc)
        r_off[p0:p0 + k] = pos + np.arange(k, dtype=np.int64) * ref_len
        pos += len(rc)
        blocks += [qc, rc]
        p0 += k
    seqs = np.concatenate(blocks)
    return PairBatch(seqs, q_off, q_len, r_off, r_len, *params, name="C2-bsj-refinement")


def concat_batches(batches, name="mixed", shuffle_seed=None):
    """One batch out of several with the same scoring parameters (config C5 style mixtures)."""
    p = (batches[0].match, batches[0].mismatch, batches[0].gap_open, batches[0].gap_extend)
    for b in batches:
        assert (b.match, b.mismatch, b.gap_open, b.gap_extend) == p, "a batch has one scoring scheme"
    base = np.concatenate([[0], np.cumsum([len(b.seqs) for b in batches])])
    seqs = np.concatenate([b.seqs for b in batches])
    q_off = np.concatenate([b.q_off + base[i] for i, b in enumerate(batches)])
    r_off = np.concatenate([b.r_off + base[i] for i, b in enumerate(batches)])
    q_len = np.concatenate([b.q_len for b in batches])
    r_len = np.concatenate([b.r_len for b in batches])
    if shuffle_seed is not None:
        perm = np.random.default_rng(shuffle_seed).permutation(len(q_len))
        q_off, r_off, q_len, r_len = q_off[perm], r_off[perm], q_len[perm], r_len[perm]
    return PairBatch(seqs, q_off, q_len, r_off, r_len, *p, name=name)


def long_query_short_ref_pairs(n_pairs, seed=SEED_BASE + 7, q_min=200, q_max=5000, ref_len=50, params=(10, 4, 8, 2)):
    """S4/S6-like pairs (collapse.py:251-256, 373-387): a 0.2-5 kb read against a 50-nt junction sequence."""
    rng = np.random.default_rng(seed)
    ql = rng.integers(q_min, q_max + 1, size=n_pairs)
    q = rng.integers(0, 4, size=int(ql.sum()), dtype=np.int8)
    q_off = np.cumsum(ql) - ql
    st = (rng.random(n_pairs) * (ql - ref_len + 1)).astype(np.int64)
    src = _ragged_arange(q_off + st, np.full(n_pairs, ref_len))
    r_codes, r_len = noisy_channel(q[src], np.full(n_pairs, ref_len, dtype=np.int64), rng)
    seqs, q_off2, r_off = _pack(q, ql.astype(np.int32), r_codes, r_len)
    return PairBatch(seqs, q_off2, ql.astype(np.int32), r_off, r_len, *params, name="S4S6-long-query")
