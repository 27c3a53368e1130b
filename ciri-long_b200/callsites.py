"""Two-phase (collect, then one device batch) versions of CIRI-long's SSW call sites.

SURVEY.md section 8(f) rank 2: the reference calls ``Aligner(ref, ...).align(query)`` once per pair from
inside per-read / per-cluster Python loops (find_bsj.py:182-233, collapse.py:161-173, 210-215).  The
functions here keep the reference's post-processing of each alignment result line for line, but take a
*list* of work items and issue one ``align_pairs`` batch for all of them.  Genome access stays with the
caller (the reference pulls windows from ``env.GENOME``, a pysam handle that is out of scope here): items
carry the window sequence they would have fetched.
"""
from collections import Counter, namedtuple

import numpy as np

from . import ssw_wrap

_COMP = str.maketrans("ATCGNatcgn", "TAGCNtagcn")

Hit = namedtuple("Hit", "q_st q_en r_st r_en strand")


def revcomp(seq):
    """CIRI_long/utils.py revcomp"""
    return seq.translate(_COMP)[::-1]


def align_clip_segments_batch(items, device=0):
    """Batched ``find_bsj.align_clip_segments`` (find_bsj.py:182-233).

    items: list of (circ, hit, window_seq, tmp_start, tmp_end) where ``window_seq`` is what the reference
    reads with ``env.GENOME.seq(hit.ctg, tmp_start, tmp_end)`` for
    ``tmp_start = max(hit.r_st - 200000, 0)``, ``tmp_end = min(hit.r_en + 200000, contig_len)``.
    ``window_seq`` may be None for items that do not reach the alignment (fewer than 20 clipped bases).
    Returns one ``(clipped_circ, circ_start, circ_end, (clip_r_st, clip_r_en, clip_base))`` per item,
    exactly what the reference returns (``(None, None, None, None)`` for its early exits)."""
    out = [None] * len(items)
    refs, queries, owners = [], [], []
    for k, (circ, hit, window, tmp_start, tmp_end) in enumerate(items):
        st_clip, en_clip = hit.q_st, len(circ) - hit.q_en
        if st_clip + en_clip >= 20:
            clip_seq = circ[hit.q_en:] + circ[:hit.q_st]
            if len(clip_seq) > 0.6 * len(circ):
                out[k] = (None, None, None, None)
                continue
            if Counter(window)['N'] >= 0.3 * (tmp_end - tmp_start):
                out[k] = (None, None, None, None)
                continue
            refs.append(window if hit.strand > 0 else revcomp(window))
            queries.append(clip_seq)
            owners.append(k)
        else:
            out[k] = (circ[hit.q_st:] + circ[:hit.q_st], hit.r_st - 1, hit.r_en, (None, None, st_clip + en_clip))
    # phase 2: one device batch with the find_bsj scoring (match=1, mismatch=1, gap_open=1, gap_extend=1);
    # only coordinates are consumed, so the CIGAR pass is skipped
    results = ssw_wrap.align_pairs(refs, queries, 1, 1, 1, 1, device=device, need_cigar=False) if owners else []
    for k, align_res in zip(owners, results):
        circ, hit, window, tmp_start, tmp_end = items[k]
        clip_seq = circ[hit.q_en:] + circ[:hit.q_st]
        if hit.strand > 0:
            clip_r_st, clip_r_en = tmp_start + align_res.ref_begin, tmp_start + align_res.ref_end
            moved = clip_r_st < hit.r_st
        else:
            clip_r_st, clip_r_en = tmp_end - align_res.ref_end, tmp_end - align_res.ref_begin
            moved = clip_r_en > hit.r_en
        if moved:
            clipped_circ = clip_seq[align_res.query_begin:] + circ[hit.q_st:hit.q_en] + clip_seq[:align_res.query_begin]
        else:
            clipped_circ = circ[hit.q_st:] + circ[:hit.q_st]
        clip_base = hit.q_st + len(circ) - hit.q_en - (align_res.query_end - align_res.query_begin) + 1
        out[k] = (clipped_circ, min(hit.r_st, clip_r_st) - 1, max(hit.r_en, clip_r_en), (clip_r_st, clip_r_en, clip_base))
    return out


def junc_score_batch(genomic_spans, junc_seq_lists, device=0):
    """Batched ``collapse.junc_score`` (collapse.py:210-215) for several candidate junctions at once.

    genomic_spans[i] is ``env.GENOME.seq(ctg, junc[0], junc[1])`` of candidate i (the reference doubles it);
    junc_seq_lists[i] the junction reads scored against it.  Returns the mean SSW score per candidate
    (collapse scoring 10/4/8/2)."""
    refs, queries, owner = [], [], []
    for i, (span, seqs) in enumerate(zip(genomic_spans, junc_seq_lists)):
        doubled = span * 2
        for s in seqs:
            refs.append(doubled); queries.append(s); owner.append(i)
    res = ssw_wrap.align_pairs(refs, queries, 10, 4, 8, 2, device=device, need_cigar=False) if refs else []
    sums = np.zeros(len(genomic_spans)); cnt = np.zeros(len(genomic_spans))
    for i, r in zip(owner, res):
        sums[i] += r.score; cnt[i] += 1
    return (sums / np.maximum(cnt, 1)).tolist()


def curate_junction_batch(candidates, junc, distance, device=0):
    """Batched ``collapse.curate_junction`` (collapse.py:161-173).

    candidates: list of (i, j, genomic_junction_seq) — the 20-nt ``genome_junction_seq(ctg, i, j, width=10)``
    of every (i, j) the reference's double loop visits; junc: the POA junction consensus; distance: the
    edit-distance function of ``utils.distance`` (edlib / Levenshtein, not vendored by the reference).
    Returns ``sorted([(i, j, avg_score)], key=score)`` like the reference."""
    from operator import itemgetter
    refs = [c[2] for c in candidates]
    res = ssw_wrap.align_pairs(refs, [junc] * len(refs), 10, 4, 8, 2, device=device, need_cigar=False) if refs else []
    scores = []
    for (i, j, tmp), alignment in zip(candidates, res):
        x = junc[alignment.query_begin:alignment.query_end]           # avg_score, collapse.py:156-158
        scores.append((i, j, distance(tmp, x) / len(tmp)))
    return sorted(scores, key=itemgetter(2))
