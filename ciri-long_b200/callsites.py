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

# CIRI_long/utils.py:118-120: upper-case A/T/C/G only -- lower-case (soft-masked) bases are reversed but NOT
# complemented by the reference, and the results must stay identical to it
_COMP = str.maketrans("ATCG", "TAGC")

Hit = namedtuple("Hit", "q_st q_en r_st r_en strand")


def _coords(refs, queries, match, mismatch, gap_open, gap_extend, device, shared_ref=False):
    """Score and coordinates of a batch as plain int arrays (no per-pair Python objects): what every call site
    below except refined_sequences_batch needs from ``Aligner.align``."""
    rec, _ = ssw_wrap.align_pairs(refs, queries, match, mismatch, gap_open, gap_extend, device=device,
                                  need_cigar=False, _shared_ref=shared_ref, as_records=True)
    # (status 1 only concerns the CIGAR -- the traceback left the band --, coordinates are exact)
    if len(rec) and ((rec["status"] & 0xff) > 1).any():
        raise ssw_wrap.SSWCudaError("alignment refused for %d pair(s)" % int(((rec["status"] & 0xff) > 1).sum()))
    return (rec["score1"].astype(np.int64), rec["ref_begin1"].astype(np.int64), rec["ref_end1"].astype(np.int64),
            rec["read_begin1"].astype(np.int64), rec["read_end1"].astype(np.int64))


def revcomp(seq):
    """CIRI_long/utils.py:118-120 (same translation table, lower-case letters pass through unchanged)"""
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
    if owners:
        _, rb, re_, qb, qe = _coords(refs, queries, 1, 1, 1, 1, device)
    for n, k in enumerate(owners):
        ref_begin, ref_end, query_begin, query_end = int(rb[n]), int(re_[n]), int(qb[n]), int(qe[n])
        circ, hit, window, tmp_start, tmp_end = items[k]
        clip_seq = circ[hit.q_en:] + circ[:hit.q_st]
        if hit.strand > 0:
            clip_r_st, clip_r_en = tmp_start + ref_begin, tmp_start + ref_end
            moved = clip_r_st < hit.r_st
        else:
            clip_r_st, clip_r_en = tmp_end - ref_end, tmp_end - ref_begin
            moved = clip_r_en > hit.r_en
        if moved:
            clipped_circ = clip_seq[query_begin:] + circ[hit.q_st:hit.q_en] + clip_seq[:query_begin]
        else:
            clipped_circ = circ[hit.q_st:] + circ[:hit.q_st]
        clip_base = hit.q_st + len(circ) - hit.q_en - (query_end - query_begin) + 1
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
    sums = np.zeros(len(genomic_spans)); cnt = np.zeros(len(genomic_spans))
    if refs:
        score = _coords(refs, queries, 10, 4, 8, 2, device)[0]
        np.add.at(sums, owner, score)
        np.add.at(cnt, owner, 1)
    return (sums / np.maximum(cnt, 1)).tolist()


def curate_junction_batch(candidates, junc, distance=None, device=0):
    """Batched ``collapse.curate_junction`` (collapse.py:161-173).

    candidates: list of (i, j, genomic_junction_seq) — the 20-nt ``genome_junction_seq(ctg, i, j, width=10)``
    of every (i, j) the reference's double loop visits; junc: the POA junction consensus; distance: a
    per-pair edit-distance function with the meaning of ``utils.distance``; by default the distances of all
    candidates are computed in one device batch (``distance.distance_batch``).
    Returns ``sorted([(i, j, avg_score)], key=score)`` like the reference."""
    from operator import itemgetter
    refs = [c[2] for c in candidates]
    if refs:
        _, _, _, qb, qe = _coords(refs, [junc] * len(refs), 10, 4, 8, 2, device)
        pieces = [junc[int(b):int(e)] for b, e in zip(qb, qe)]                          # avg_score, collapse.py:156-158
    else:
        pieces = []
    if distance is None:
        from .distance import distance_batch
        dists = distance_batch(refs, pieces, device) if refs else []
    else:
        dists = [distance(tmp, x) for tmp, x in zip(refs, pieces)]
    scores = [(i, j, int(d) / len(tmp)) for (i, j, tmp), d in zip(candidates, dists)]
    return sorted(scores, key=itemgetter(2))


# ---- helpers the call sites below post-process with (restated; the reference keeps them in utils.py / align.py)
def transform_seq(seq, bsj):
    """CIRI_long/utils.py:123-124"""
    return seq[bsj:] + seq[:bsj]


def get_junc_seq(seq, bsj, width=25):
    """CIRI_long/utils.py:127-140"""
    st, en = bsj - width, bsj + width
    if len(seq) <= 2 * width:
        return seq[bsj - len(seq) // 2:] + seq[:bsj - len(seq) // 2]
    if st < 0:
        if en < 0:
            return seq[st:en]
        return seq[st:] + seq[:en]
    if en > len(seq):
        return seq[st:] + seq[:en - len(seq)]
    return seq[st:en]


_CIGAR_OPS = {c: k for k, c in enumerate("MIDNSHP=X")}


def find_alignment_pos(alignment, pos):
    """CIRI_long/align.py:803-820: query position aligned to reference position ``pos`` (None if outside)."""
    import re
    r_st = r_en = alignment.ref_begin
    q_st = q_en = alignment.query_begin
    for l, op in re.findall(r'(\d+)([MIDNSHP=X])', alignment.cigar_string):
        l, op = int(l), _CIGAR_OPS[op]
        if op == 0:
            r_en += l; q_en += l
        elif op == 1:
            q_en += l
        elif op == 2:
            r_en += l
        if r_st <= pos <= r_en:
            return q_st + pos - r_st
        r_st, q_st = r_en, q_en
    return None


def cluster_junction_seqs_batch(clusters, device=0):
    """Batched head of ``collapse.correct_cluster`` (collapse.py:251-265) for several clusters at once.

    clusters: list of ``(ref_seq, [query_seq, ...])`` -- the sequence of the longest 'full' read of the most
    common circ_id, and the sequences of ``cluster[1:]`` (at least one, as the reference's ``max(head_pos)``
    requires).  Two device batches instead of two per-read loops per cluster: every read against the first
    50 nt of its cluster's reference (head position), then every read against the rotated template.
    Returns one ``(template, junc_seqs)`` per cluster, ``junc_seqs`` being the list handed to ``spoa.poa``."""
    refs, queries, owner = [], [], []
    for c, (ref_seq, qs) in enumerate(clusters):
        if not qs:
            raise ValueError("cluster %d has no query reads (the reference fails on max([]) here)" % c)
        for q in qs:
            refs.append(ref_seq[:50]); queries.append(q); owner.append(c)
    ref_begin = _coords(refs, queries, 10, 4, 8, 2, device)[1]
    head = [[] for _ in clusters]
    for c, rb in zip(owner, ref_begin):
        head[c].append(int(rb))
    templates = [transform_seq(ref_seq, max(h)) for (ref_seq, _), h in zip(clusters, head)]
    query_begin = _coords([templates[c] for c in owner], queries, 10, 4, 8, 2, device)[3]
    out = [(t, [get_junc_seq(t, -max(h) // 2, 25)]) for t, h in zip(templates, head)]
    for c, q, qb in zip(owner, queries, query_begin):
        out[c][1].append(get_junc_seq(transform_seq(q, int(qb)), -max(head[c]) // 2, 25))
    return out


def refined_sequences_batch(items, device=0):
    """Batched rotation of the 'full' reads onto the curated junction (collapse.py:369-385).

    items: list of ``(circ_junc_seq, [(read_id, seq), ...])`` -- ``genome_junction_seq(ctg, circ_start,
    circ_end)`` and the (already sampled and sorted) full reads of the cluster.  This is the one CIRI-long
    call site that reads the CIGAR.  Returns ``cluster_seq`` per item: ``[(read_id, rotated_seq), ...]``."""
    refs, queries, owner = [], [], []
    for c, (junc, reads) in enumerate(items):
        for k, (_, seq) in enumerate(reads):
            refs.append(junc); queries.append(seq * 2); owner.append((c, k))
    res = ssw_wrap.align_pairs(refs, queries, 10, 4, 8, 2, report_cigar=True, device=device) if refs else []
    out = [[None] * len(reads) for _, reads in items]
    for (c, k), alignment in zip(owner, res):
        junc, reads = items[c]
        read_id, seq = reads[k]
        # no CIGAR (alignment refused, or its traceback left the band and the reference's CIGAR is undefined):
        # the read keeps its rotation, like a junction position outside the alignment
        tmp_pos = None if alignment is None or alignment.cigar_string is None else find_alignment_pos(alignment, len(junc) // 2)
        out[c][k] = (read_id, seq) if tmp_pos is None else (read_id, transform_seq(seq, tmp_pos % len(seq)))
    return out


def exon_scores_batch(consensus_seq, exon_pair_seqs, device=0):
    """Batched ``collapse.exon_score`` (collapse.py:760-774) for the candidates of one ``iter_flow`` step
    (collapse.py:730,750,755): every candidate's concatenated (and, on the minus strand, reverse-complemented)
    exon-pair sequence against the isoform consensus.  Returns ``ref_end - ref_begin`` per candidate."""
    if not exon_pair_seqs:
        return []
    _, rb, re_, _, _ = _coords([consensus_seq] * len(exon_pair_seqs), exon_pair_seqs, 10, 4, 8, 2, device, shared_ref=True)
    return [int(e - b) for b, e in zip(rb, re_)]
