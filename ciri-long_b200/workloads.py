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


# ---- generators on the GPU (torch): the same recipes as above, a different random stream; a million pairs take
# ---- seconds instead of minutes.  All return host (numpy) PairBatches.
def _t_ragged_arange(starts, lens):
    import torch
    lens = lens.long()
    total = int(lens.sum())
    seg_start = torch.cumsum(lens, 0) - lens
    return torch.repeat_interleave(starts.long() - seg_start, lens) + torch.arange(total, device=lens.device)


def noisy_channel_torch(codes, seg_len, g, sub=0.05, ins=0.04, dele=0.04, max_run=3, n_frac=0.0):
    """noisy_channel on the device: (codes int8 tensor, seg_len long tensor) -> (codes, seg_len)."""
    import torch
    dev = codes.device
    total = codes.numel()
    seg_id = torch.repeat_interleave(torch.arange(len(seg_len), device=dev), seg_len.long())
    u = torch.rand(total, device=dev, generator=g)
    keep = u >= dele
    is_sub = keep & (u < dele + sub)
    shift = torch.randint(1, 4, (total,), dtype=torch.int8, device=dev, generator=g)
    codes = torch.where(is_sub, (codes + shift) % 4, codes)
    runs = torch.where(torch.rand(total, device=dev, generator=g) < ins,
                       torch.randint(1, max_run + 1, (total,), device=dev, generator=g),
                       torch.zeros((), dtype=torch.long, device=dev))
    reps = keep.long() + runs
    new = torch.repeat_interleave(codes, reps)
    grp_start = torch.cumsum(reps, 0) - reps
    pos = torch.arange(new.numel(), device=dev) - torch.repeat_interleave(grp_start, reps)
    inserted = pos >= torch.repeat_interleave(keep.long(), reps)
    rnd = torch.randint(0, 4, (new.numel(),), dtype=torch.int8, device=dev, generator=g)
    new = torch.where(inserted, rnd, new)
    if n_frac > 0:
        new = torch.where(torch.rand(new.numel(), device=dev, generator=g) < n_frac,
                          torch.full((), 4, dtype=torch.int8, device=dev), new)
    new_len = torch.bincount(torch.repeat_interleave(seg_id, reps), minlength=len(seg_len))
    return new, new_len


def _pack_torch(q_codes, q_len, r_codes, r_len):
    """[q0 r0 q1 r1 ...] in one buffer, on the device; returns host arrays (seqs, q_off, r_off)."""
    import torch
    n = len(q_len)
    lens = torch.stack([q_len.long(), r_len.long()], 1).reshape(-1)
    offs = torch.cumsum(lens, 0) - lens
    q_off, r_off = offs[0::2], offs[1::2]
    seqs = torch.empty(int(lens.sum()), dtype=torch.int8, device=q_codes.device)
    seqs[_t_ragged_arange(q_off, q_len)] = q_codes
    seqs[_t_ragged_arange(r_off, r_len)] = r_codes
    return seqs.cpu().numpy(), q_off.cpu().numpy().astype(np.int64), r_off.cpu().numpy().astype(np.int64)


def square_pairs_torch(n_pairs, length, device, seed=SEED_BASE + 4, params=(1, 1, 1, 1)):
    """Config C4 on the device: m ~ n ~ length."""
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed + length)
    refs = torch.randint(0, 4, (n_pairs * length,), dtype=torch.int8, device=device, generator=g)
    ql0 = torch.full((n_pairs,), length, dtype=torch.long, device=device)
    q, ql = noisy_channel_torch(refs, ql0, g)
    seqs, q_off, r_off = _pack_torch(q, ql, refs, ql0)
    return PairBatch(seqs, q_off, ql.cpu().numpy().astype(np.int32), r_off, np.full(n_pairs, length, dtype=np.int32), *params,
                     name="C4-square-%d" % length)


def rolling_circle_pairs_torch(n_reads, device, seed=SEED_BASE + 3, read_min=2000, read_max=6000, copies_min=2, copies_max=8,
                               params=(10, 4, 8, 2), chunk=20000):
    """Config C3 on the device (rolling_circle_pairs recipe), generated in chunks of reads."""
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    parts = []
    for c0 in range(0, n_reads, chunk):
        n = min(chunk, n_reads - c0)
        read_len = torch.randint(read_min, read_max + 1, (n,), device=device, generator=g)
        copies = torch.randint(copies_min, copies_max + 1, (n,), device=device, generator=g)
        unit_len = torch.clamp(read_len // copies, min=20)
        seg_unit = torch.repeat_interleave(torch.arange(n, device=device), copies)
        seg_len0 = unit_len[seg_unit]
        unit_off = torch.cumsum(unit_len, 0) - unit_len
        units = torch.randint(0, 4, (int(unit_len.sum()),), dtype=torch.int8, device=device, generator=g)
        rot = (torch.rand(n, device=device, generator=g) * unit_len.float()).long()
        idx_in = _t_ragged_arange(torch.zeros(len(seg_len0), dtype=torch.long, device=device), seg_len0)
        seg_rep = torch.repeat_interleave(torch.arange(len(seg_len0), device=device), seg_len0)
        su = seg_unit[seg_rep]
        src = unit_off[su] + (idx_in + rot[su]) % unit_len[su]
        seg_codes, seg_len = noisy_channel_torch(units[src], seg_len0, g)
        seg_off = torch.cumsum(seg_len, 0) - seg_len
        first_seg = torch.cumsum(copies, 0) - copies
        is_query = torch.ones(len(seg_len), dtype=torch.bool, device=device)
        is_query[first_seg] = False
        q_idx = torch.nonzero(is_query).squeeze(1)
        r_idx = first_seg[seg_unit[q_idx]]
        ok = (seg_len[q_idx] > 0) & (seg_len[r_idx] > 0)
        q_idx, r_idx = q_idx[ok], r_idx[ok]
        parts.append(PairBatch(seg_codes.cpu().numpy(), seg_off[q_idx].cpu().numpy().astype(np.int64),
                               seg_len[q_idx].cpu().numpy().astype(np.int32), seg_off[r_idx].cpu().numpy().astype(np.int64),
                               seg_len[r_idx].cpu().numpy().astype(np.int32), *params, name="C3-rolling-circle"))
    return concat_batches(parts, name="C3-rolling-circle") if len(parts) > 1 else parts[0]


def junction_pairs_torch(n_pairs, device, seed=SEED_BASE + 5, q_min=40, q_max=60, ref_len=20, params=(10, 4, 8, 2)):
    """S2-like tiny pairs on the device (junction_pairs recipe)."""
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    ql = torch.randint(q_min, q_max + 1, (n_pairs,), device=device, generator=g)
    q = torch.randint(0, 4, (int(ql.sum()),), dtype=torch.int8, device=device, generator=g)
    q_off = torch.cumsum(ql, 0) - ql
    st = (torch.rand(n_pairs, device=device, generator=g) * (ql - ref_len + 1).float()).long()
    rl0 = torch.full((n_pairs,), ref_len, dtype=torch.long, device=device)
    src = _t_ragged_arange(q_off + st, rl0)
    r_codes, r_len = noisy_channel_torch(q[src], rl0, g)
    seqs, q_off2, r_off = _pack_torch(q, ql, r_codes, r_len)
    return PairBatch(seqs, q_off2, ql.cpu().numpy().astype(np.int32), r_off, r_len.cpu().numpy().astype(np.int32), *params,
                     name="S2-junction")


def long_query_short_ref_pairs_torch(n_pairs, device, seed=SEED_BASE + 7, q_min=200, q_max=5000, ref_len=50, params=(10, 4, 8, 2)):
    """S4/S6-like pairs on the device (long_query_short_ref_pairs recipe)."""
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    ql = torch.randint(q_min, q_max + 1, (n_pairs,), device=device, generator=g)
    q = torch.randint(0, 4, (int(ql.sum()),), dtype=torch.int8, device=device, generator=g)
    q_off = torch.cumsum(ql, 0) - ql
    st = (torch.rand(n_pairs, device=device, generator=g) * (ql - ref_len + 1).float()).long()
    rl0 = torch.full((n_pairs,), ref_len, dtype=torch.long, device=device)
    src = _t_ragged_arange(q_off + st, rl0)
    r_codes, r_len = noisy_channel_torch(q[src], rl0, g)
    seqs, q_off2, r_off = _pack_torch(q, ql, r_codes, r_len)
    return PairBatch(seqs, q_off2, ql.cpu().numpy().astype(np.int32), r_off, r_len.cpu().numpy().astype(np.int32), *params,
                     name="S4S6-long-query")


def clip_window_pairs_torch(n_pairs, device, seed=SEED_BASE + 8, q_min=20, q_max=600, win_min=400000, win_max=600000,
                            genome_len=1 << 26, params=(1, 1, 1, 1)):
    """S1 (find_bsj.py:191-215): the clipped bases of a read (20-600 nt) against the +-200 kb genomic window around
    its hit.  Windows are views into one synthetic genome (offsets are explicit in the C ABI, so the windows of
    neighbouring reads overlap in memory exactly like the reference's `env.GENOME.seq(ctg, st, en)` slices of one
    chromosome); the clip is a noisy copy of a stretch inside its window."""
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    genome = torch.randint(0, 4, (genome_len,), dtype=torch.int8, device=device, generator=g)
    wl = torch.randint(win_min, win_max + 1, (n_pairs,), device=device, generator=g)
    ws = (torch.rand(n_pairs, device=device, generator=g, dtype=torch.float64) * (genome_len - wl).double()).long()
    ql0 = torch.randint(q_min, q_max + 1, (n_pairs,), device=device, generator=g)
    qs = ws + (torch.rand(n_pairs, device=device, generator=g, dtype=torch.float64) * (wl - ql0).double()).long()
    src = _t_ragged_arange(qs, ql0)
    q, ql = noisy_channel_torch(genome[src], ql0, g, n_frac=0.005)
    ql = torch.clamp(ql, min=1)
    q_off = genome_len + torch.cumsum(ql, 0) - ql
    seqs = torch.cat([genome, q[:int(ql.sum())]]).cpu().numpy()
    return PairBatch(seqs, q_off.cpu().numpy().astype(np.int64), ql.cpu().numpy().astype(np.int32), ws.cpu().numpy().astype(np.int64),
                     wl.cpu().numpy().astype(np.int32), *params, name="S1-clip-vs-window")


def repack(batch, order=None):
    """Rebuild the code buffer in pair order [q0 r0 q1 r1 ...] (optionally after permuting the pairs), so that any
    contiguous run of pairs references a compact byte range -- what the chunked one-shot call uploads."""
    n = len(batch)
    order = np.arange(n) if order is None else np.asarray(order)
    ql, rl = batch.q_len[order], batch.r_len[order]
    q_src = _ragged_arange(batch.q_off[order], ql)
    r_src = _ragged_arange(batch.r_off[order], rl)
    seqs, q_off, r_off = _pack(batch.seqs[q_src], ql, batch.seqs[r_src], rl)
    return PairBatch(seqs, q_off, ql.copy(), r_off, rl.copy(), batch.match, batch.mismatch, batch.gap_open, batch.gap_extend,
                     batch.name)


def mixed_slab_torch(n_pairs, device, seed, params=(10, 4, 8, 2)):
    """One slab of config C5: 60 % S2-like junction pairs (collapse.py:165-172), 25 % C2-like
    segment-vs-flank pairs, 10 % C3-like rolling-circle segment pairs (collapse.py:259-265), 5 % S4/S6-like long
    reads vs a 50-nt junction (collapse.py:373-387); one scoring scheme (collapse's 10/4/8/2), shuffled, repacked
    pair-major."""
    n2 = int(n_pairs * 0.60); nc2 = int(n_pairs * 0.25); n46 = int(n_pairs * 0.05)
    nc3 = n_pairs - n2 - nc2 - n46
    parts = [junction_pairs_torch(n2, device, seed=seed + 1, params=params),
             bsj_refinement_pairs_torch(nc2, device, seed=seed + 2, params=params),
             long_query_short_ref_pairs_torch(n46, device, seed=seed + 4, params=params)]
    c3 = rolling_circle_pairs_torch(max(8, int(nc3 / 3.6)), device, seed=seed + 3, params=params)
    if len(c3) >= nc3:
        c3 = PairBatch(c3.seqs, c3.q_off[:nc3], c3.q_len[:nc3], c3.r_off[:nc3], c3.r_len[:nc3], *params, name=c3.name)
    parts.append(c3)
    mix = concat_batches(parts, name="C5-mixed")
    order = np.random.default_rng(seed).permutation(len(mix))
    out = repack(mix, order)
    out.name = "C5-mixed"
    return out
