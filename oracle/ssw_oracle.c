/*
 * ssw_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A scalar, CPU-only restatement of the striped Smith-Waterman semantics of the
 * CIRI-long reference (libs/striped_smith_waterman/ssw.c).  It is the checker for
 * the CUDA path; nothing under ciri-long_b200/ may call, link or import it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity status: PINNED.  tests/test_oracle.py compares this file against
 *   (a) the golden vectors in tests/golden/ (generated from the unmodified reference
 *       libssw.so by oracle/make_golden.py) and
 *   (b) the live reference build in oracle/_ref/libssw.so when it is present.
 *
 * The restatement is deliberately *not* SIMD code: every 128-bit register of the
 * reference becomes an explicit array of L lanes (L = 16 for the byte kernel, 8 for
 * the word kernel) and every saturating instruction becomes a scalar helper, so that
 * the layout-dependent behaviour of the reference (query rows padded to L*segLen,
 * lazy-F termination rules, second-best scan, band bookkeeping of the CIGAR pass) is
 * reproduced bit for bit while the text stays readable.
 *
 * Reference map (file:line in /root/reference/libs/striped_smith_waterman/):
 *   orc_score_pass(mode=BYTE)  <- ssw.c:89-114 (profile) + ssw.c:123-345 (sw_sse2_byte)
 *   orc_score_pass(mode=WORD)  <- ssw.c:347-369 (profile) + ssw.c:371-546 (sw_sse2_word)
 *   orc_band_cigar             <- ssw.c:548-735 (banded_sw)
 *   orc_align                  <- ssw.c:750-771 (ssw_init: bias) + ssw.c:779-869 (ssw_align)
 *   orc_cigar_op/len           <- ssw.c:876-902
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ssw_oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* lane-vector helpers: one "register" = L unsigned 16-bit slots (only 8 bits used in BYTE mode) */

#define ORC_MAXL 16

typedef struct { int32_t v[ORC_MAXL]; } lanes_t;

static inline int32_t sat_sub_unsigned(int32_t a, int32_t b) { return a > b ? a - b : 0; }

static inline int32_t sat_add_mode(int32_t a, int32_t b, int word)
{
    int32_t s = a + b;
    if (word) {                      /* signed 16-bit saturating add (ssw.c:442) */
        if (s > 32767) s = 32767;
        if (s < -32768) s = -32768;
    } else {                         /* unsigned 8-bit saturating add (ssw.c:205) */
        if (s > 255) s = 255;
        if (s < 0) s = 0;
    }
    return s;
}

/* shift every lane up by one slot, lane 0 receives 0 (ssw.c:195,248,264,429,469) */
static inline void lanes_shift_up(lanes_t* x, int L)
{
    for (int l = L - 1; l > 0; --l) x->v[l] = x->v[l - 1];
    x->v[0] = 0;
}

static inline int32_t lanes_hmax(const lanes_t* x, int L)
{
    int32_t m = x->v[0];
    for (int l = 1; l < L; ++l) if (x->v[l] > m) m = x->v[l];
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* One score pass (forward or reverse) in the byte or the word flavour.
 *
 * The query profile is not materialised: prof(nt, j, l) is evaluated on the fly as
 *   row = j + l*segLen;  row >= readLen ? pad : mat[nt*n + read[row]] (+ bias in BYTE mode)
 * which is exactly what ssw.c:104-112 / 359-367 tabulate.
 */
void orc_score_pass(int word, const int8_t* ref, int ref_dir, int32_t refLen,
                    const int8_t* read, int32_t readLen, const int8_t* mat, int32_t n,
                    uint8_t gapO, uint8_t gapE, int32_t terminate, uint8_t bias, int32_t maskLen,
                    orc_ends* out)
{
    const int L = word ? 8 : 16;
    const int32_t segLen = (readLen + L - 1) / L;
    const int32_t useBias = word ? 0 : bias;
    int32_t max = 0;
    int32_t end_read = readLen - 1;
    int32_t end_ref = word ? 0 : -1;                       /* ssw.c:145 vs ssw.c:388 */

    int32_t* maxColumn = (int32_t*)calloc(refLen > 0 ? refLen : 1, sizeof(int32_t));
    lanes_t* Hstore = (lanes_t*)calloc(segLen > 0 ? segLen : 1, sizeof(lanes_t));
    lanes_t* Hload  = (lanes_t*)calloc(segLen > 0 ? segLen : 1, sizeof(lanes_t));
    lanes_t* E      = (lanes_t*)calloc(segLen > 0 ? segLen : 1, sizeof(lanes_t));
    lanes_t* Hmax   = (lanes_t*)calloc(segLen > 0 ? segLen : 1, sizeof(lanes_t));
    lanes_t vMaxScore, vMaxMark;
    memset(&vMaxScore, 0, sizeof vMaxScore);
    memset(&vMaxMark, 0, sizeof vMaxMark);

    int32_t begin = 0, end = refLen, step = 1;
    if (ref_dir == 1) { begin = refLen - 1; end = -1; step = -1; }

    for (int32_t i = begin; i != end; i += step) {
        lanes_t vF, vMaxColumn, vH;
        memset(&vF, 0, sizeof vF);
        memset(&vMaxColumn, 0, sizeof vMaxColumn);
        vH = Hstore[segLen - 1];
        lanes_shift_up(&vH, L);
        const int32_t nt = ref[i];
        { lanes_t* t = Hload; Hload = Hstore; Hstore = t; }

        /* main sweep over the segLen vector slots (ssw.c:204-238 / 441-465) */
        for (int32_t j = 0; j < segLen; ++j) {
            for (int l = 0; l < L; ++l) {
                const int32_t row = j + l * segLen;
                int32_t p;
                if (word) p = row >= readLen ? 0 : mat[nt * n + read[row]];
                else      p = row >= readLen ? bias : (int32_t)(uint8_t)(int8_t)(mat[nt * n + read[row]] + bias);
                int32_t h = sat_add_mode(vH.v[l], p, word);
                if (!word) h = sat_sub_unsigned(h, useBias);
                int32_t e = E[j].v[l];
                if (e > h) h = e;
                if (vF.v[l] > h) h = vF.v[l];
                if (h > vMaxColumn.v[l]) vMaxColumn.v[l] = h;
                Hstore[j].v[l] = h;
                h = sat_sub_unsigned(h, gapO);
                e = sat_sub_unsigned(e, gapE);
                if (h > e) e = h;
                E[j].v[l] = e;
                int32_t f = sat_sub_unsigned(vF.v[l], gapE);
                vF.v[l] = f > h ? f : h;
                vH.v[l] = Hload[j].v[l];
            }
        }

        if (!word) {
            /* byte flavour lazy-F (ssw.c:240-272): test first, unbounded wrap-around */
            int32_t j = 0;
            vH = Hstore[0];
            lanes_shift_up(&vF, L);
            for (;;) {
                int any = 0;
                for (int l = 0; l < L; ++l)
                    if (sat_sub_unsigned(vF.v[l], sat_sub_unsigned(vH.v[l], gapO)) != 0) any = 1;
                if (!any) break;
                for (int l = 0; l < L; ++l) {
                    if (vF.v[l] > vH.v[l]) vH.v[l] = vF.v[l];
                    if (vH.v[l] > vMaxColumn.v[l]) vMaxColumn.v[l] = vH.v[l];
                    vF.v[l] = sat_sub_unsigned(vF.v[l], gapE);
                }
                Hstore[j] = vH;
                ++j;
                if (j >= segLen) { j = 0; lanes_shift_up(&vF, L); }
                vH = Hstore[j];
            }
        } else {
            /* word flavour lazy-F (ssw.c:467-478): update first, test after, at most 8 wraps,
             * vMaxColumn is NOT refreshed here. */
            int done = 0;
            for (int k = 0; k < 8 && !done; ++k) {
                lanes_shift_up(&vF, L);
                for (int32_t j = 0; j < segLen; ++j) {
                    int any = 0;
                    for (int l = 0; l < L; ++l) {
                        int32_t h = Hstore[j].v[l];
                        if (vF.v[l] > h) h = vF.v[l];
                        Hstore[j].v[l] = h;
                        h = sat_sub_unsigned(h, gapO);
                        vF.v[l] = sat_sub_unsigned(vF.v[l], gapE);
                        if (vF.v[l] > h) any = 1;
                    }
                    if (!any) { done = 1; break; }
                }
            }
        }

        /* running maximum (ssw.c:274-291 / 481-495) */
        int changed = 0;
        for (int l = 0; l < L; ++l) {
            if (vMaxColumn.v[l] > vMaxScore.v[l]) vMaxScore.v[l] = vMaxColumn.v[l];
            if (vMaxMark.v[l] != vMaxScore.v[l]) changed = 1;
        }
        if (changed) {
            vMaxMark = vMaxScore;
            int32_t temp = lanes_hmax(&vMaxScore, L);
            if (temp > max) {
                max = temp;
                if (!word && max + bias >= 255) break;           /* overflow: ssw.c:285 */
                end_ref = i;
                for (int32_t j = 0; j < segLen; ++j) Hmax[j] = Hstore[j];
            }
        }
        maxColumn[i] = lanes_hmax(&vMaxColumn, L);
        if (maxColumn[i] == terminate) break;
    }

    /* alignment end on the read: smallest row holding the maximum (ssw.c:299-308 / 502-511) */
    for (int32_t j = 0; j < segLen; ++j)
        for (int l = 0; l < L; ++l)
            if (Hmax[j].v[l] == max) {
                const int32_t row = j + l * segLen;
                if (row < end_read) end_read = row;
            }

    out->score = (!word && max + bias >= 255) ? 255 : max;
    out->ref = end_ref;
    out->read = end_read;
    out->score2 = 0;
    out->ref2 = 0;

    /* second best outside the mask window (ssw.c:325-340 / 528-541) */
    int32_t edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;
    for (int32_t i = 0; i < edge; ++i)
        if (maxColumn[i] > out->score2) { out->score2 = maxColumn[i]; out->ref2 = i; }
    edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
    for (int32_t i = word ? edge : edge + 1; i < refLen; ++i)
        if (maxColumn[i] > out->score2) { out->score2 = maxColumn[i]; out->ref2 = i; }

    free(maxColumn); free(Hstore); free(Hload); free(E); free(Hmax);
}

/* ------------------------------------------------------------------------------------------ */
/* Banded affine DP + traceback over the trimmed rectangle (ssw.c:548-735).
 *
 * Band rows are stored in "line" coordinates: for matrix cell (i, j) and band half-width w,
 *   off(i) = max(i - w, 0),  slot(i, j) = j - off(i) + 1      (the reference's set_u)
 * hb = previous row H, eb = vertical-gap scores (shared between the two rows), hc = current row H.
 * Direction codes per cell: [0] vertical-gap source (2 extend / 3 open), [1] horizontal-gap
 * source (4 extend / 5 open), [2] H source (1 diagonal, else a copy of [0] or [1]).
 *
 * Defined deviation: cells the reference never writes hold 0 here, so a traceback that leaves the
 * band ends in the reference's own "Trace back error" outcome (status ORC_ERR_TRACEBACK) instead of
 * reading stale heap.
 */
static inline int32_t band_off(int32_t i, int32_t w) { int32_t x = i - w; return x > 0 ? x : 0; }

int orc_band_cigar(const int8_t* ref, const int8_t* read, int32_t refLen, int32_t readLen,
                   int32_t score, uint32_t gapO, uint32_t gapE, int32_t band_width,
                   const int8_t* mat, int32_t n, uint32_t** cigar_out, int32_t* cigar_len_out,
                   int32_t* final_band_out)
{
    const int32_t go = (int32_t)gapO, ge = (int32_t)gapE;
    int32_t max = 0;
    int32_t width = 0, width_d = 0;
    int32_t* hb = NULL; int32_t* eb = NULL; int32_t* hc = NULL;
    int8_t* dir = NULL;
    size_t line_cap = 0, dir_cap = 0;

    do {
        width = band_width * 2 + 3;
        width_d = band_width * 2 + 1;
        if ((size_t)width + 1 > line_cap) {
            size_t nc = (size_t)width + 1;
            hb = (int32_t*)realloc(hb, nc * sizeof(int32_t));
            eb = (int32_t*)realloc(eb, nc * sizeof(int32_t));
            hc = (int32_t*)realloc(hc, nc * sizeof(int32_t));
            /* keep earlier contents (the reference never resets eb between bands), zero the growth */
            for (size_t k = line_cap; k < nc; ++k) hb[k] = eb[k] = hc[k] = 0;
            line_cap = nc;
        }
        size_t need = (size_t)width_d * (size_t)readLen * 3 + 1;
        if (need > dir_cap) { dir = (int8_t*)realloc(dir, need); dir_cap = need; }
        memset(dir, 0, dir_cap);

        for (int32_t j = 1; j < width - 1; ++j) hb[j] = 0;
        for (int32_t i = 0; i < readLen; ++i) {
            int32_t beg = i - band_width; if (beg < 0) beg = 0;
            int32_t end = i + band_width; if (end > refLen - 1) end = refLen - 1;
            const int32_t edge = end + 1 < width - 1 ? end + 1 : width - 1;
            int32_t f = 0, u = 0;
            hb[0] = eb[0] = hb[edge] = eb[edge] = hc[0] = 0;
            int8_t* line = dir + (size_t)width_d * i * 3;
            const int32_t off_cur = band_off(i, band_width), off_up = band_off(i - 1, band_width);

            for (int32_t j = beg; j <= end; ++j) {
                u = j - off_cur + 1;
                const int32_t up = j - off_up + 1;         /* (i-1, j)   */
                const int32_t left = u - 1;                /* (i, j-1)   */
                const int32_t diag = j - 1 - off_up + 1;   /* (i-1, j-1) */
                int8_t* cell = line + (size_t)(j - off_cur) * 3;

                int32_t open = (i == 0 ? 0 : hb[up]) - go;
                int32_t ext  = (i == 0 ? 0 : eb[up]) - ge;
                eb[u] = open > ext ? open : ext;
                cell[0] = open > ext ? 3 : 2;

                open = hc[left] - go;
                ext = f - ge;
                f = open > ext ? open : ext;
                cell[1] = open > ext ? 5 : 4;

                const int32_t e1 = eb[u] > 0 ? eb[u] : 0;
                const int32_t f1 = f > 0 ? f : 0;
                const int32_t gapbest = e1 > f1 ? e1 : f1;
                const int32_t dg = hb[diag] + mat[ref[j] * n + read[i]];
                hc[u] = gapbest > dg ? gapbest : dg;
                if (hc[u] > max) max = hc[u];
                if (gapbest <= dg) cell[2] = 1;
                else cell[2] = e1 > f1 ? cell[0] : cell[1];
            }
            for (int32_t j = 1; j <= u; ++j) hb[j] = hc[j];
        }
        band_width *= 2;
    } while (max < score && band_width < 2 * readLen);
    band_width /= 2;
    if (final_band_out) *final_band_out = band_width;

    /* traceback from the bottom-right corner in state H, stops at read row 0 (ssw.c:636-696) */
    int32_t cap = 16, l = 0, run = 0;
    uint32_t* c = (uint32_t*)malloc(cap * sizeof(uint32_t));
    int32_t i = readLen - 1, j = refLen - 1, state = 2;
    char op = 'M', prev_op = 'M';
    int status = ORC_OK;
    while (i > 0) {
        const int32_t col = j - band_off(i, band_width);
        int8_t d = 0;
        if (col >= 0 && col < width_d) d = dir[(size_t)width_d * i * 3 + (size_t)col * 3 + state];
        switch (d) {
            case 1: --i; --j; state = 2; op = 'M'; break;
            case 2: --i; state = 0; op = 'I'; break;
            case 3: --i; state = 2; op = 'I'; break;
            case 4: --j; state = 1; op = 'D'; break;
            case 5: --j; state = 2; op = 'D'; break;
            default: status = ORC_ERR_TRACEBACK; break;
        }
        if (status != ORC_OK) break;
        if (op == prev_op) ++run;
        else {
            if (l + 2 >= cap) { cap *= 2; c = (uint32_t*)realloc(c, cap * sizeof(uint32_t)); }
            c[l++] = orc_to_cigar_int((uint32_t)run, prev_op);
            prev_op = op;
            run = 1;
        }
    }
    free(hb); free(eb); free(hc); free(dir);
    if (status != ORC_OK) { free(c); *cigar_out = NULL; *cigar_len_out = 0; return status; }

    if (l + 3 >= cap) { cap += 4; c = (uint32_t*)realloc(c, cap * sizeof(uint32_t)); }
    if (op == 'M') c[l++] = orc_to_cigar_int((uint32_t)run + 1, op);     /* ssw.c:697-704 */
    else { c[l++] = orc_to_cigar_int((uint32_t)run, op); c[l++] = orc_to_cigar_int(1, 'M'); }

    uint32_t* r = (uint32_t*)malloc((size_t)l * sizeof(uint32_t));
    for (int32_t k = 0; k < l; ++k) r[k] = c[l - 1 - k];
    free(c);
    *cigar_out = r;
    *cigar_len_out = l;
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
int orc_align(const int8_t* read, int32_t readLen, const int8_t* ref, int32_t refLen,
              const int8_t* mat, int32_t n, int8_t score_size,
              uint8_t gapO, uint8_t gapE, uint8_t flag, uint16_t filters, int32_t filterd,
              int32_t maskLen, orc_result* r)
{
    memset(r, 0, sizeof *r);
    r->ref_begin1 = -1;
    r->read_begin1 = -1;

    const int have_byte = (score_size == 0 || score_size == 2);
    const int have_word = (score_size == 1 || score_size == 2);
    int32_t bias = 0;
    if (have_byte) {                                           /* ssw.c:756-762 */
        for (int32_t k = 0; k < n * n; ++k) if (mat[k] < bias) bias = mat[k];
        bias = -bias;
    }

    orc_ends fwd;
    int word = 0;
    if (have_byte) {
        orc_score_pass(0, ref, 0, refLen, read, readLen, mat, n, gapO, gapE, 255, (uint8_t)bias, maskLen, &fwd);
        if (fwd.score == 255) {
            if (!have_word) return ORC_ERR_OVERFLOW;           /* ssw.c:810-813 */
            orc_score_pass(1, ref, 0, refLen, read, readLen, mat, n, gapO, gapE, 65535, 0, maskLen, &fwd);
            word = 1;
        }
    } else if (have_word) {
        orc_score_pass(1, ref, 0, refLen, read, readLen, mat, n, gapO, gapE, 65535, 0, maskLen, &fwd);
        word = 1;
    } else return ORC_ERR_NOPROFILE;

    r->word = word;
    r->score1 = (uint16_t)fwd.score;
    r->ref_end1 = fwd.ref;
    r->read_end1 = fwd.read;
    if (maskLen >= 15) { r->score2 = (uint16_t)fwd.score2; r->ref_end2 = fwd.ref2; }
    else { r->score2 = 0; r->ref_end2 = -1; }
    if (flag == 0 || (flag == 2 && r->score1 < filters)) return ORC_OK;

    /* begin coordinates: same flavour on the reversed read prefix, ref walked right-to-left */
    const int32_t plen = r->read_end1 + 1;
    int8_t* rev = (int8_t*)calloc(plen > 0 ? plen : 1, 1);
    for (int32_t k = 0; k < plen; ++k) rev[k] = read[plen - 1 - k];
    orc_ends bwd;
    orc_score_pass(word, ref, 1, r->ref_end1 + 1, rev, plen, mat, n, gapO, gapE, r->score1,
                   (uint8_t)bias, maskLen, &bwd);
    free(rev);
    r->ref_begin1 = bwd.ref;
    r->read_begin1 = r->read_end1 - bwd.read;

    if ((7 & flag) == 0 || ((2 & flag) != 0 && r->score1 < filters) ||
        ((4 & flag) != 0 && (r->ref_end1 - r->ref_begin1 > filterd || r->read_end1 - r->read_begin1 > filterd)))
        return ORC_OK;

    if (r->ref_begin1 < 0) {
        /* score 0 in the byte flavour: the reference goes on to read ref[-1] (undefined), but the
         * rectangle is 1x1 and the traceback loop never runs, so the CIGAR is always "1M". */
        r->cigar = (uint32_t*)malloc(sizeof(uint32_t));
        r->cigar[0] = orc_to_cigar_int(1, 'M');
        r->cigarLen = 1;
        r->band_width = 1;
        return ORC_OK;
    }
    const int32_t subRef = r->ref_end1 - r->ref_begin1 + 1;
    const int32_t subRead = r->read_end1 - r->read_begin1 + 1;
    int32_t bw = subRef - subRead; if (bw < 0) bw = -bw; bw += 1;
    return orc_band_cigar(ref + r->ref_begin1, read + r->read_begin1, subRef, subRead, r->score1,
                          gapO, gapE, bw, mat, n, &r->cigar, &r->cigarLen, &r->band_width);
}

void orc_result_free(orc_result* r) { free(r->cigar); r->cigar = NULL; r->cigarLen = 0; }

char orc_cigar_op(uint32_t c)
{
    static const char map[] = "MIDNSHP=X";
    const uint32_t code = c & 0xfU;
    return code < 9 ? map[code] : 'M';
}
uint32_t orc_cigar_len(uint32_t c) { return c >> 4; }

/* ------------------------------------------------------------------------------------------ */
/* Batch driver used by tests and by bench.py's cpu_baseline "port" leg: struct-of-arrays in,
 * flat result records + a concatenated CIGAR buffer out.  Single-threaded on purpose. */
int orc_align_batch(int32_t n_pairs, const int8_t* seqs,
                    const int64_t* q_off, const int32_t* q_len,
                    const int64_t* r_off, const int32_t* r_len,
                    const int8_t* mat, int32_t n, uint8_t gapO, uint8_t gapE, uint8_t flag,
                    const int32_t* mask_len,
                    orc_flat* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used)
{
    int64_t used = 0;
    for (int32_t p = 0; p < n_pairs; ++p) {
        orc_result r;
        int st = orc_align(seqs + q_off[p], q_len[p], seqs + r_off[p], r_len[p], mat, n, 2,
                           gapO, gapE, flag, 0, 0, mask_len[p], &r);
        orc_flat* o = out + p;
        o->status = st; o->word = r.word; o->band_width = r.band_width;
        o->score1 = r.score1; o->score2 = r.score2;
        o->ref_begin1 = r.ref_begin1; o->ref_end1 = r.ref_end1;
        o->read_begin1 = r.read_begin1; o->read_end1 = r.read_end1; o->ref_end2 = r.ref_end2;
        o->cigar_off = used; o->cigar_len = 0;
        if (st == ORC_OK && r.cigarLen > 0) {
            if (used + r.cigarLen > cigar_cap) { orc_result_free(&r); return ORC_ERR_CAPACITY; }
            memcpy(cigar_buf + used, r.cigar, (size_t)r.cigarLen * sizeof(uint32_t));
            o->cigar_len = r.cigarLen;
            used += r.cigarLen;
        }
        orc_result_free(&r);
    }
    if (cigar_used) *cigar_used = used;
    return ORC_OK;
}
