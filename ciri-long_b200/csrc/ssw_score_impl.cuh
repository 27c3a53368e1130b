#pragma once
// ssw_score_impl.cuh -- score passes of the striped Smith-Waterman hot path for sm_100a.
//
// Replaces sw_sse2_byte / sw_sse2_word (reference ssw.c:123-345, 371-546) for the forward pass and
// for the reverse pass of ssw_align (ssw.c:836-849).  Not a port of the SSE2 code:
//
//   * inter-task parallelism, one warp per (query, reference) pair;
//   * the query is cut into "virtual strips" of up to K consecutive rows: lane t owns strip t in the low
//     16-bit half and strip t+32 in the high half of every register, so each DPX instruction
//     (VIADDMNMX.S16x2 / VIMNMX3.S16x2) updates two cells;
//   * a systolic wavefront over the reference: at step s strip v computes column s-v; H, F and the
//     running column maximum leave a strip through one rotate-shuffle per value and step;
//   * substitution scores come from a lane-replicated (bank == lane, conflict free) shared-memory
//     table indexed by (ref base pair, query base pair): one LDS per two cells, off the DPX pipe;
//   * all scores are kept biased by +gap_open (H' = H + go, E' = E + go, F' = F + go): the floor of the
//     local alignment becomes max(., go) inside a VIMNMX3, and H - gap_open is a plain 32-bit subtract of
//     two halves that can never borrow, so it leaves the DPX pipe: 4 DPX instructions per two cells;
//   * queries longer than one tile of 64 strips are processed in row tiles that hand H/F/colmax to the
//     next tile through a per-warp boundary array.
//
// What is computed is the *semantics* of the reference, verified against oracle/ssw_oracle.c:
//   GOTOH  plain affine-gap recurrences on the real query rows (equal to the reference's byte flavour,
//          and to its word flavour when gap_open > gap_extend), the flavour (8/16 bit) being decided
//          after the pass from max+bias >= 255 exactly like ssw.c:285,317,806.  The reference pads the
//          query to a multiple of 16 (8) rows with zero-scoring rows that only influence maxColumn[];
//          their contribution is added in closed form from the last real row (second_best()).
//   TRUNC  the word flavour when gap_open == gap_extend: the reference's lazy-F loop stops after one
//          step (ssw.c:467-478), so the vertical-gap chain is cut at every segment boundary
//          (row % ceil(m/8) == 0) and only the boundary row's H sees the incoming F.  Strips are aligned
//          to the reference's 8 segments, so the cut always falls on the first row of a strip and only
//          that row pays for it; a strip holds K or K-1 rows of its segment (the unused last row is
//          skipped by selecting the hand-off from row K-2), and the reference's pad rows are simply the
//          tail of the last segment.
#include <stdio.h>
#include <atomic>
#include <type_traits>
#include "ssw_common.cuh"
#include "ssw_kernels.h"
#include "ssw_second_best.cuh"

namespace sswb {

// Row layout of one strip (see header): first query row, number of rows it holds, segment start flag.
struct StripGeom { int first, live; bool valid, segStart; };

template <int K, bool TRUNC>
__device__ __forceinline__ StripGeom strip_geom(int v, int Vtot, int dead, int segLen, int G, int base, int extra)
{
    StripGeom s;
    if (!TRUNC) {
        s.first = v * K - dead;            // may be negative: zero rows in front of row 0
        s.live = K; s.valid = true; s.segStart = false;
    } else {
        s.valid = v < Vtot;
        const int l = v / G, g = v - l * G;
        s.live = s.valid ? base + (g < extra ? 1 : 0) : 0;
        s.first = l * segLen + g * base + (g < extra ? g : extra);
        s.segStart = s.valid && g == 0 && l >= 1;
    }
    return s;
}

// A strip's H column at the step of a new maximum, K words per (half, lane), written as 16-byte vectors:
// the store is executed by the whole warp for one active lane, so fewer, wider stores are cheaper.
template <int K>
__device__ __forceinline__ void store_snapshot(unsigned* dst, const unsigned (&Hd)[K])
{
    constexpr int KP = (K + 3) & ~3;
#pragma unroll
    for (int i = 0; i < KP; i += 4) {
        uint4 v;
        v.x = Hd[i];
        v.y = i + 1 < K ? Hd[i + 1] : 0u;
        v.z = i + 2 < K ? Hd[i + 2] : 0u;
        v.w = i + 3 < K ? Hd[i + 3] : 0u;
        *reinterpret_cast<uint4*>(dst + i) = v;
    }
}

template <int K, bool TRUNC, bool REV, bool CHUNK>
__device__ __forceinline__ void score_pair(const ScoreArgs& a, const int pair, const unsigned* lut, unsigned char* rpw, unsigned char* ws,
                                           const int c0 = -1, const int c1 = 0, const int task = 0)
{
    const unsigned FULL = 0xffffffffu;
    constexpr int KP = (K + 3) & ~3;                                // words per snapshot
    const int lane = lane_id();
    PairRec* rec = a.b.rec + pair;

    int m, n, terminate = 0;
    const int8_t* qb;
    const int8_t* rb;
    int qs, rs;
    if (!REV) {
        m = a.b.q_len[pair];
        n = a.b.r_len[pair];
        qb = a.b.seqs + a.b.q_off[pair];
        rb = a.b.seqs + a.b.r_off[pair];
        qs = 1; rs = 1;
    } else {
        // reversed read prefix [0, read_end1] against ref[0, ref_end1] walked right-to-left (ssw.c:837-844)
        m = rec->read_end1 + 1;
        n = rec->ref_end1 + 1;
        qb = a.b.seqs + a.b.q_off[pair] + rec->read_end1;
        rb = a.b.seqs + a.b.r_off[pair] + rec->ref_end1;
        qs = -1; rs = -1;
        terminate = rec->score1;
    }
    const int go = a.sc.go, ge = a.sc.ge;

    // per-warp scratch: [column records | tile boundary | best-column snapshots]
    unsigned* colbuf = reinterpret_cast<unsigned*>(ws + a.off_col);
    uint2* bnd = reinterpret_cast<uint2*>(ws + a.off_bnd);
    unsigned* snap = reinterpret_cast<unsigned*>(ws + a.off_snap);

    // ---- strip layout
    int T, Vtot, dead = 0, segLen = 0, G = 1, base = 0, extra = 0;
    if (!TRUNC) {
        const int rpt = VSTRIPS * K;
        T = (m + rpt - 1) / rpt;
        dead = T * rpt - m;
        Vtot = T * VSTRIPS;
    } else {
        segLen = (m + 7) / 8;                                       // ssw.c:389
        G = segLen > 8 * KMAX ? (segLen + KMAX - 1) / KMAX : 8;     // strips per segment
        base = segLen / G; extra = segLen - base * G;
        Vtot = 8 * G;
        T = (Vtot + VSTRIPS - 1) / VSTRIPS;
        if (base + (extra > 0 ? 1 : 0) != K) {                      // list construction and kernel disagree: never guess
            if (lane == 0) rec->status |= PS_PUNT;
            return;
        }
    }

    // ---- forward pass over one column chunk of a long reference (ssw_kernels.h: ChunkPlan): the wavefront
    // runs over columns [cw, c1) of the reference; columns before c0 = cw + skip only warm the state up.
    // The few values this needs later (skip, cw, whether the task is the whole pair) wait in the spare bytes
    // of the warp's shared-memory window instead of registers that the inner loop has no room for.
    int* const ckw = reinterpret_cast<int*>(rpw + RP_WINDOW - 12);
    bool limited = false;                                           // reverse pass: first look at a bounded number of columns
    if (CHUNK) {
        int cw = 0, skip = 0;
        if (c0 >= 0) {
            if (c0 > 0) {
                // the writer strip (63: chunks are cut only for single-tile queries) reaches column `skip` at step
                // skip + 63; skip is a multiple of RP_CHUNK so that this is a border of the sweep's blocks
                const int ov = chunk_overlap(m, a.ck.max_match, ge);
                skip = ((ov + RP_CHUNK - 1) / RP_CHUNK) * RP_CHUNK;
                cw = c0 - skip;
                if (cw <= 0) { cw = 0; skip = 0; }                  // exact from column 0: duplicates of the first chunk's work are harmless
            }
            rb += (long long)cw * rs;
            const int whole = (c0 == 0 && c1 == n) ? 1 : 0;
            n = c1 - cw;
            if (lane == 0) { ckw[0] = skip; ckw[1] = cw; ckw[2] = whole; }
        } else {
            if (lane == 0) { ckw[0] = 0; ckw[1] = 0; ckw[2] = 1; }
            if (REV) {
                // Reverse pass of a long reference, first look: the stop column (ssw.c:296,499) normally lies about one
                // alignment length away.  Only if it is not found within rev_look(m) columns is the pair expanded
                // into column-chunk tasks over the whole prefix.
                const int look = rev_look(m);
                if (n > look && chunk_tasks(m, n, a.ck.chunk_cols, a.ck.max_match, ge) > 1) { n = look; limited = true; }
            }
        }
        __syncwarp();
    }

    const unsigned GO = pack2(go, go);                              // bias of every stored score, and the local floor
    const unsigned mge = pack2(-ge, -ge);
    const int src = (lane + 31) & 31;
    // lane 0: high half <- lane 31's low half, low half <- second operand (the bias for H/F, 0 for colmax)
    const unsigned fix = lane == 0 ? 0x1054u : 0x3210u;
    const unsigned lutS = (unsigned)__cvta_generic_to_shared(lut) + lane * 4;     // shared-window address of my LUT column
    const unsigned rpS = (unsigned)__cvta_generic_to_shared(rpw) + 31 - lane;      // my slot of the ref-pair window at t = 0
    const unsigned mtermP = pack2(-(terminate + go), -(terminate + go));

    int candM = 0, candCol = -1, candRow = 0;
    int termCol = -1, overCol = 0x7fffffff;
    // Reverse pass only, three ways to run it:
    //   fast     (single-tile queries): no best-cell tracking at all.  The answer is the first column whose maximum
    //            equals score1 (ssw.c:296,499) and the first row in it that holds score1, so every strip only notes
    //            the first column in which one of its rows reaches score1 exactly (and the row); cells above
    //            score1 are noted like below.  If no column reaches score1 the tracking way decides.
    //   tracking plain best-cell tracking, cells above score1 kept out of it and only their first column noted.
    //   exact    if a cell above score1 turns up at or before the stop column (the two truncated-F passes
    //            disagree): limited to the columns the reference visits, plain best-cell tracking.
    constexpr int RM_FAST = 0, RM_TRACK = 1, RM_EXACT = 2;
    // (not for the truncated-F recurrence: its reverse pass often stays below score1 -- the segment borders fall on
    // other rows than in the forward pass --, and a fast pass that finds nothing has to be repeated with tracking)
    int mode = (REV && !CHUNK && !TRUNC && T == 1) ? RM_FAST : RM_TRACK;
    bool exactMode = false;
    const unsigned mterm1P = pack2(1 - (terminate + go), 1 - (terminate + go));
    int hitColLo = -1, hitColHi = -1, hitRowLo = 0, hitRowHi = 0;

    for (;;) {
    exactMode = mode == RM_EXACT;
    candM = 0; candCol = -1; candRow = 0; termCol = -1; overCol = 0x7fffffff;
    hitColLo = -1; hitColHi = -1;
    for (int p = 0; p < T; ++p) {
        const bool lastTile = (p == T - 1);
        const StripGeom sLo = strip_geom<K, TRUNC>(p * VSTRIPS + lane, Vtot, dead, segLen, G, base, extra);
        const StripGeom sHi = strip_geom<K, TRUNC>(p * VSTRIPS + lane + 32, Vtot, dead, segLen, G, base, extra);
        unsigned qoff[K], E[K], Hd[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int rl = sLo.first + i, rh = sHi.first + i;
            int cl = 4, ch = 4;
            if (rl >= 0 && rl < m && i < sLo.live) { cl = qb[(long long)rl * qs]; if ((unsigned)cl > 4u) cl = 4; }
            if (rh >= 0 && rh < m && i < sHi.live) { ch = qb[(long long)rh * qs]; if ((unsigned)ch > 4u) ch = 4; }
            qoff[i] = lutS + (unsigned)(cl * 5 + ch) * 128u;                // LUT address of this row pair for ref pair 0
            E[i] = GO; Hd[i] = GO;
        }
        // TRUNC only: which halves take their hand-off from row K-1 (strip holds K rows) and which from row
        // K-2 (K-1 rows); gates of the first row; strips past the end of the query do not count
        const unsigned selMask = (sLo.live == K ? 0xffffu : 0u) | (sHi.live == K ? 0xffff0000u : 0u);
        const unsigned stripMask = (sLo.valid ? 0xffffu : 0u) | (sHi.valid ? 0xffff0000u : 0u);
        const unsigned g0 = pack2(sLo.segStart ? TRUNC_GATE : 0, sHi.segStart ? TRUNC_GATE : 0);
        const unsigned gF0 = pack2(sLo.segStart ? TRUNC_GATE : -ge, sHi.segStart ? TRUNC_GATE : -ge);
        // the strip whose hand-off is the complete column: colmax / H of the last row leave the tile there
        const int wv = (TRUNC && lastTile) ? Vtot - 1 - p * VSTRIPS : VSTRIPS - 1;
        const int wLane = wv & 31, wHalf = wv >> 5;

        unsigned Hout = GO, Fout = GO, R = 0, diagIn = GO, best = GO;
        // best  = per strip, the largest H recorded with its column and a snapshot of the strip's H column
        // bestT = max(best, the largest H any strip of this warp has recorded, refreshed every 32 steps):
        //         a cell below that can never be the pair's maximum, so it is not recorded at all
        unsigned bestT = GO;
        int bcolLo = -1, bcolHi = -1;
        int cLo = -lane, cHi = -lane - 32;
        int termflag = 0;
        const int steps = n + wv;
        const bool isWriter = lane == wLane;
        const unsigned selW = wHalf ? 0x7632u : 0x5410u;                // (colmax, H last row) of the writer's half
        unsigned* wcol = colbuf + (wHalf ? cHi : cLo);                  // the writer's column record at step 0
        uint2* wbnd = bnd + cHi;
        unsigned rpCur = 24u;                                           // ref-pair code of the current step (Z,Z before column 0)

        // One wavefront step at offset t of the current chunk.
        //   CHECK = some lane may be outside [0, n) (pipeline fill and drain)
        //   MULTI = the query spans several tiles: strip 0 may continue below the previous tile (boundary
        //           array in), the last strip may feed the next tile (boundary array out)
        auto step = [&](auto chk, auto multi, auto fastc, const int s, const int t) {
            constexpr bool CHECK = decltype(chk)::value;
            constexpr bool MULTI = decltype(multi)::value;
            constexpr bool FAST = REV && decltype(fastc)::value;
            // hand-off from the previous strip (computed one step ago, same column as ours now)
            unsigned rH = __byte_perm(__shfl_sync(FULL, Hout, src), GO, fix);
            unsigned rF = __byte_perm(__shfl_sync(FULL, Fout, src), GO, fix);
            unsigned rR = __byte_perm(__shfl_sync(FULL, R, src), 0u, fix);
            if (MULTI) {
                if (p > 0 && lane == 0 && s < n) {
                    const uint2 bv = bnd[s];
                    rH = (rH & 0xffff0000u) | (bv.x & 0xffffu);
                    rF = (rF & 0xffff0000u) | (bv.x >> 16);
                    rR = (rR & 0xffff0000u) | (bv.y & 0xffffu);
                }
            }
            unsigned diag = diagIn;
            diagIn = rH;
            unsigned F = rF;
            const unsigned rp = rpCur;
            rpCur = lds_u8(rpS + t + 1);                                  // next step's code, one step ahead

            unsigned mx = 0, hprev = 0, Hk2 = rH, Fk2 = rF;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const unsigned sc = lds_u32(mad_u32(rp, RP_STRIDE, qoff[i]));
                const unsigned x = addmax(diag, sc, E[i]);                    // max(Hdiag + s, E)
                const unsigned h = max3(x, F, GO);                            // H (floor 0 == bias)
                unsigned u;
                if (TRUNC && i == 0) {
                    // first row of a strip: if it starts a segment the incoming F reaches H only
                    const unsigned tt = addmax(F, g0, x);
                    const unsigned h0 = max3(tt, GO, GO);
                    u = h0 - GO;
                    F = addmax(F, gF0, u);
                } else {
                    u = h - GO;                                               // H - gap_open: halves cannot borrow
                    F = addmax(F, mge, u);
                }
                E[i] = addmax(E[i], mge, u);
                diag = Hd[i];
                Hd[i] = h;
                if (TRUNC && i == K - 2) { Hk2 = h; Fk2 = F; }
                const unsigned hm = (TRUNC && i == K - 1) ? (h & selMask) : h;  // an unused last row does not count
                if (i & 1) mx = max3(mx, hprev, hm);
                hprev = hm;
            }
            if (K & 1) mx = max_relu(mx, hprev);
            if (TRUNC) {
                Hout = (Hd[K - 1] & selMask) | (Hk2 & ~selMask);
                Fout = (F & selMask) | (Fk2 & ~selMask);
            } else {
                Hout = Hd[K - 1];
                Fout = F;
            }

            unsigned mxv = mx;
            if (CHECK) {
                const unsigned vm = ((unsigned)cLo < (unsigned)n ? 0xffffu : 0u) | ((unsigned)cHi < (unsigned)n ? 0xffff0000u : 0u);
                mxv &= vm;
            }
            if (TRUNC) mxv &= stripMask;
            R = max_relu(rR, mxv);
            if (FAST) {
                const unsigned reach = addmax_relu(mxv, mterm1P, 0u);          // max(mxv - score1 + 1, 0) per half
                if (reach) {
                    const int vl = lo16(mxv) - go, vh = hi16(mxv) - go;
                    if (vl > terminate) overCol = cLo < overCol ? cLo : overCol;
                    else if (vl == terminate && hitColLo < 0) {
                        hitColLo = cLo;
#pragma unroll
                        for (int i = K - 1; i >= 0; --i) if (i < sLo.live && lo16(Hd[i]) - go == terminate) hitRowLo = sLo.first + i;
                    }
                    if (vh > terminate) overCol = cHi < overCol ? cHi : overCol;
                    else if (vh == terminate && hitColHi < 0) {
                        hitColHi = cHi;
#pragma unroll
                        for (int i = K - 1; i >= 0; --i) if (i < sHi.live && hi16(Hd[i]) - go == terminate) hitRowHi = sHi.first + i;
                    }
                }
            }
            if (REV && !FAST && !exactMode) {
                // The reference stops at the first column whose maximum equals score1 (ssw.c:296,499), so
                // cells above score1 only count if they occur before that column.  Strips ahead of the
                // stop column keep running here: values above score1 are kept out of the best-cell
                // tracking and only their first column is remembered (checked after the pass).
                // One packed op says whether any half is above the target; which half is then decided on the scalar
                // values.  (Testing the halves of `ov` itself miscompiles with ptxas 12.9: it takes both conditions
                // from the predicate outputs of the VIMNMX, one predicate register for two halves, and a lane whose
                // other half is an unused strip keeps its over-the-target cell -- found by tools/fuzz_gpu.py on
                // multi-tile queries.)
                const unsigned ov = addmax_relu(mxv, mtermP, 0u);              // max(mxv - score1, 0) per half
                if (ov) {
                    const int skipc = CHUNK ? ckw[0] : 0;               // (chunk mode: warm-up columns do not count)
                    if (lo16(mxv) - go > terminate) { if (!CHUNK || cLo >= skipc) overCol = cLo < overCol ? cLo : overCol; mxv &= 0xffff0000u; }
                    if (hi16(mxv) - go > terminate) { if (!CHUNK || cHi >= skipc) overCol = cHi < overCol ? cHi : overCol; mxv &= 0x0000ffffu; }
                }
            }
            bool pHi = true, pLo = true;
            unsigned nb = 0;
            if (!FAST) nb = __vibmax_s16x2(bestT, mxv, &pHi, &pLo);      // pred = (bestT >= mxv)
            if (!FAST && !(pHi && pLo)) {
                // (chunk mode: warm-up columns, c < skip, are not recorded and do not raise the threshold)
                const int skip = CHUNK ? ckw[0] : 0;
                unsigned acc = 0;
                if (!pLo && (!CHUNK || cLo >= skip)) {
                    best = (best & 0xffff0000u) | (mxv & 0xffffu);
                    bcolLo = cLo;
                    acc = 0xffffu;
                    store_snapshot<K>(snap + lane * KP, Hd);
                }
                if (!pHi && (!CHUNK || cHi >= skip)) {
                    best = (best & 0xffffu) | (mxv & 0xffff0000u);
                    bcolHi = cHi;
                    acc |= 0xffff0000u;
                    store_snapshot<K>(snap + (32 + lane) * KP, Hd);
                }
                if (!CHUNK) bestT = nb;
                else bestT = (nb & acc) | (bestT & ~acc);
            }
            // the complete column leaves the tile at the writer strip
            const unsigned wval = __byte_perm(R, Hout, selW);               // colmax | H(last row) << 16 of the writer's half
            bool colOk = isWriter;
            if (CHECK) colOk = colOk && (unsigned)(wHalf ? cHi : cLo) < (unsigned)n;
            if (MULTI && !lastTile) {
                if (colOk) wbnd[s] = make_uint2(__byte_perm(Hout, Fout, 0x7632u), R >> 16);   // always strip 63: high halves
            } else if (!REV) {
                if (colOk) wcol[s] = wval;
            } else if (colOk && !termflag && !exactMode && (int)(wval & 0xffffu) == terminate + go) {
                if (!CHUNK || (wHalf ? cHi : cLo) >= ckw[0]) { termflag = 1; termCol = wHalf ? cHi : cLo; }
            }
            ++cLo; ++cHi;
        };

        // Steps run in chunks of RP_CHUNK: the warp first stages the ref-pair codes of the chunk
        // (code(c) * 5 + code(c - 32), Z outside [0, n)) in its shared-memory window, then sweeps it.
        // Chunk borders fall on the phase borders: pipeline fill [0, 63), steady state [63, n), drain [n, steps).
        const int sA = steps < 63 ? steps : 63;
        int sB = n < steps ? n : steps; if (sB < sA) sB = sA;
        auto refresh = [&]() {
            const int lo = lo16(best), hi = hi16(best);
            const int g = __reduce_max_sync(FULL, lo > hi ? lo : hi);
            // g - 1: a cell that only equals the warp-wide maximum still counts if its column is earlier (ties go to
            // the first column, ssw.c:286-292), and a strip that is ahead in time can be behind in columns
            bestT = max_relu(bestT, pack2(g - 1, g - 1));
        };
        auto sweep = [&](auto multi, auto fastc) {
            constexpr bool FASTS = REV && decltype(fastc)::value;
            int s0 = 0;
            bool stop = false;
            while (s0 < steps && !stop) {
                int s1 = s0 + RP_CHUNK < steps ? s0 + RP_CHUNK : steps;
                if (s0 < sA) { if (s1 > sA) s1 = sA; }
                else if (s0 < sB) { if (s1 > sB) s1 = sB; }
                if (CHUNK && !REV) {
                    // a block must not straddle the step at which the writer strip leaves the warm-up columns (it does
                    // when the chunk's own part is shorter than the pipeline: the last chunk of a reference)
                    const int sw = ckw[0] ? ckw[0] + wv : 0;
                    if (s0 < sw && s1 > sw) s1 = sw;
                }
                const bool steady = s0 >= sA && s1 <= sB;
                if (CHUNK && !REV) {
                    // column records of the chunk's own columns go to the pair's array, warm-up columns to scratch
                    const int skip = ckw[0];
                    const bool own = skip == 0 || s0 >= skip + wv;
                    wcol = (own ? a.ck.col_pool + a.ck.col_off[pair] + ckw[1] : colbuf) - wv;
                }
                __syncwarp();
                for (int x = lane; x < s1 - s0 + 32; x += 32) {
                    const int c = s0 - 31 + x, c2 = c - 32;
                    int ca = 4, cb = 4;
                    if (c >= 0 && c < n) { ca = rb[(long long)c * rs]; if ((unsigned)ca > 4u) ca = 4; }
                    if (c2 >= 0 && c2 < n) { cb = rb[(long long)c2 * rs]; if ((unsigned)cb > 4u) cb = 4; }
                    rpw[x] = (unsigned char)(ca * 5 + cb);
                }
                __syncwarp();
                rpCur = lds_u8(rpS);                                        // code of step s0 for my column
                const int len = s1 - s0;
                if (steady) {
#pragma unroll 8
                    for (int t = 0; t < len; ++t) {
                        step(std::false_type{}, multi, fastc, s0 + t, t);
                        if (REV && lastTile && (t & 7) == 7 && __any_sync(FULL, termflag)) { stop = true; break; }
                        if (!FASTS && (t & 31) == 31) refresh();
                    }
                } else {
                    for (int t = 0; t < len; ++t) {
                        step(std::true_type{}, multi, fastc, s0 + t, t);
                        if (REV && lastTile && (t & 7) == 7 && __any_sync(FULL, termflag)) { stop = true; break; }
                    }
                }
                s0 = s1;
            }
        };
        if (T > 1) sweep(std::true_type{}, std::false_type{});
        else if (REV && mode == RM_FAST) sweep(std::false_type{}, std::true_type{});
        else sweep(std::false_type{}, std::false_type{});

        // ---- tile epilogue: best cell of this tile in reference order (max, first column, first row)
        const int vlo = lo16(best) - go, vhi = hi16(best) - go;
        const int M = __reduce_max_sync(FULL, vlo > vhi ? vlo : vhi);
        if (M > 0) {
            const int clo = vlo == M ? bcolLo : 0x7fffffff, chi = vhi == M ? bcolHi : 0x7fffffff;
            const int col = __reduce_min_sync(FULL, clo < chi ? clo : chi);
            const int st = (vlo == M && bcolLo == col) ? lane : ((vhi == M && bcolHi == col) ? lane + 32 : 1000);
            const int strip = __reduce_min_sync(FULL, st);
            const int owner = strip & 31, half = strip >> 5;
            int row = 0;
            if (lane == owner) {
                const StripGeom sg = half ? sHi : sLo;
                for (int i = sg.live - 1; i >= 0; --i) {
                    const unsigned v = snap[(half * 32 + lane) * KP + i];
                    if ((half ? hi16(v) : lo16(v)) - go == M) row = sg.first + i;
                }
                if (row > m - 1) row = m - 1;                           // pad rows never lower end_read (ssw.c:144,306)
            }
            row = __shfl_sync(FULL, row, owner);
            if (M > candM || (M == candM && col < candCol)) { candM = M; candCol = col; candRow = row; }
        }
        if (lastTile) termCol = __shfl_sync(FULL, termCol, wLane);
        __syncwarp();      // boundary array / snapshots written by this tile are read by the next one
    }
    if (CHUNK && REV && limited && termCol < 0) {
        if (lane == 0) a.next_idx[*a.next_base + atomicAdd(a.next_count, 1)] = pair;
        return;
    }
    if (REV && !exactMode) {
        overCol = __reduce_min_sync(FULL, overCol);
        if (overCol != 0x7fffffff && (termCol < 0 || overCol <= termCol)) {
            mode = RM_EXACT;
            if (termCol >= 0) n = termCol + 1;
            continue;
        }
    }
    if (REV && mode == RM_FAST) {
        // the first row that holds score1 in the stop column, over the strips that met score1 there
        const int rl = (termCol >= 0 && hitColLo == termCol) ? hitRowLo : 0x7fffffff;
        const int rh = (termCol >= 0 && hitColHi == termCol) ? hitRowHi : 0x7fffffff;
        const int row = __reduce_min_sync(FULL, rl < rh ? rl : rh);
        if (row == 0x7fffffff) { mode = RM_TRACK; continue; }             // score1 never reached: plain tracking decides
        candM = terminate; candCol = termCol; candRow = row > m - 1 ? m - 1 : row;
    }
    break;
    }

    if (CHUNK && !REV) {
        if (candM > 0) candCol += ckw[1];
        if (!ckw[2]) {
            // merge with the pair's other chunks: largest score, then first column (one task owns a column, so
            // the row comes with it); the task that finishes last carries on with the pair's epilogue
            unsigned long long key = candM > 0 ? ((unsigned long long)candM << 49) |
                                                 ((unsigned long long)(0xfffffff - candCol) << 21) | (unsigned long long)candRow : 0ull;
            int left = 0;
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                if (key) atomicMax(a.ck.pair_key + pair, key);
                __threadfence();
                left = atomicSub(a.ck.pair_left + pair, 1) - 1;
            }
            left = __shfl_sync(FULL, left, 0);
            if (left > 0) return;
            __threadfence();
            key = *reinterpret_cast<volatile unsigned long long*>(a.ck.pair_key + pair);
            candM = (int)(key >> 49);
            candCol = candM > 0 ? 0xfffffff - (int)((key >> 21) & 0xfffffffull) : -1;
            candRow = candM > 0 ? (int)(key & 0x1fffffull) : 0;
        }
        n = a.b.r_len[pair];
        colbuf = a.ck.col_pool + a.ck.col_off[pair];
    }

    if (CHUNK && REV && c0 >= 0 && !ckw[2]) {
        // reverse pass in column chunks: every task leaves (its stop column, its best cell up to there); the task
        // that finishes last walks them in scan order up to the first stop column (ssw.c:296,499)
        const int cw = ckw[1];
        if (lane == 0)
            a.ck.task_res[task] = make_int4(termCol >= 0 ? termCol + cw : -1, candM, candM > 0 ? candCol + cw : -1, candRow);
        int left = 0;
        __threadfence();
        __syncwarp();
        if (lane == 0) left = atomicSub(a.ck.pair_left + pair, 1) - 1;
        left = __shfl_sync(FULL, left, 0);
        if (left > 0) return;
        __threadfence();
        if (lane == 0) {
            const unsigned long long span = *reinterpret_cast<volatile unsigned long long*>(a.ck.pair_key + pair);
            const int t0 = (int)(span >> 32), nt = (int)(span & 0xffffffffull);
            candM = 0; candCol = -1; candRow = 0;
            for (int t = t0; t < t0 + nt; ++t) {
                const int4 r = __ldcg(a.ck.task_res + t);          // written by other SMs: read through L2
                if (r.y > candM) { candM = r.y; candCol = r.z; candRow = r.w; }
                if (r.x >= 0) break;
            }
        }
        candM = __shfl_sync(FULL, candM, 0); candCol = __shfl_sync(FULL, candCol, 0); candRow = __shfl_sync(FULL, candRow, 0);
    }

    if (!REV) {
        const bool over8 = candM + a.sc.bias >= 255;               // ssw.c:285,317
        int word = TRUNC ? 1 : (over8 ? 1 : 0);
        int status = 0;
        if (!TRUNC && a.rerun) {
            // second look at a pair whose truncated pass stayed below the 8-bit limit: the byte flavour is
            // authoritative unless it overflows, in which case the word result already stored stands.
            if (over8) { if (lane == 0) rec->status &= ~PS_NEED_GOTOH; return; }
        } else if (!TRUNC && over8 && go == ge) {
            // the byte flavour overflows and the word flavour is the truncated-F recurrence: hand the pair over
            if (lane == 0) {
                const int pos = atomicAdd(a.next_count, 1);
                a.next_idx[*a.next_base + pos] = pair;
            }
            return;
        }
        if (TRUNC && !over8 && !a.rerun) status |= PS_NEED_GOTOH;   // rerun = the byte pass already overflowed
        // near the range where the reference's 16-bit saturation (or this kernel's gate constant) matters:
        // the pair is re-done by the 32-bit kernel (ssw_score32.cu)
        const bool wide = candM >= (TRUNC ? TRUNC_SCORE_LIMIT : S16_SCORE_LIMIT) - go;
        if (wide) status |= PS_WIDE32;

        const int endRef = candM > 0 ? candCol : (word ? 0 : -1);  // ssw.c:145 vs ssw.c:388
        const int endRead = candM > 0 ? candRow : 0;
        int score2 = 0, ref2 = -1;
        const int maskLen = a.b.mask_len[pair];
        if (maskLen >= 15) {                                        // ssw.c:826-832
            ref2 = 0;
            // TRUNC computed the reference's pad rows itself; GOTOH adds them here for the flavour that won
            const int L = word ? 8 : 16;
            const int P = TRUNC ? 0 : ((m + L - 1) / L) * L - m;
            second_best(colbuf, n, P, go, word, endRef, maskLen, go, ge, lane, score2, ref2);
        }
        if (lane == 0) {
            rec->score1 = candM; rec->score2 = score2;
            rec->ref_begin1 = -1; rec->ref_end1 = endRef;
            rec->read_begin1 = -1; rec->read_end1 = endRead;
            rec->ref_end2 = ref2; rec->cigar_len = 0; rec->cigar_off = 0;
            rec->word = word;
            rec->status = status;
            if (TRUNC && (status & PS_NEED_GOTOH)) {
                const int pos = atomicAdd(a.next_count, 1);
                a.next_idx[*a.next_base + pos] = pair;
            }
            if (wide) {
                const int pos = atomicAdd(a.wide_count, 1);
                a.wide_idx[pos] = pair;
            }
        }
    } else {
        int status = 0;
        if (candM >= (TRUNC ? TRUNC_SCORE_LIMIT : S16_SCORE_LIMIT) - go) status |= PS_PUNT;
        if (lane == 0) {
            const int word = rec->word;
            rec->ref_begin1 = candM > 0 ? rec->ref_end1 - candCol : (word ? 0 : -1);
            rec->read_begin1 = rec->read_end1 - (candM > 0 ? candRow : 0);
            rec->status |= status;
        }
    }
}

template <int K, bool TRUNC, bool REV, bool CHUNK>
__global__ void __launch_bounds__(score_warps(K) * 32, 1) score_kernel(const ScoreArgs a)
{
    extern __shared__ unsigned lut[];
    const int count = *a.wl.count;
    if (count <= 0) return;
    for (int e = threadIdx.x; e < LUT_ENTRIES * 32; e += blockDim.x) {
        const int entry = e >> 5;
        const int rp = entry / 25, qp = entry - rp * 25;
        const int rl = rp / 5, rh = rp - rl * 5, ql = qp / 5, qh = qp - ql * 5;
        const int sl = (rl == 4 || ql == 4) ? 0 : a.sc.mat[rl * 5 + ql];
        const int sh = (rh == 4 || qh == 4) ? 0 : a.sc.mat[rh * 5 + qh];
        lut[e] = pack2(sl, sh);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int base = a.wl.base ? *a.wl.base : 0;
    unsigned char* rpw = reinterpret_cast<unsigned char*>(lut) + LUT_BYTES + warp * RP_WINDOW;
    unsigned char* ws = a.scratch + (size_t)(blockIdx.x * score_warps(K) + warp) * a.scratch_stride;
    for (;;) {
        int idx = 0;
        if (lane_id() == 0) idx = atomicAdd(a.wl.cursor, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        int pair, c0 = -1, c1 = 0;
        if (CHUNK && a.ck.task_pair) { pair = a.ck.task_pair[idx]; c0 = a.ck.task_c0[idx]; c1 = a.ck.task_c1[idx]; }
        else pair = a.wl.idx[base + idx];
        score_pair<K, TRUNC, REV, CHUNK>(a, pair, lut, rpw, ws, c0, c1, idx);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// host-side launch table

template <int K, bool TRUNC, bool REV, bool CHUNK>
static cudaError_t launch_one(const ScoreArgs& a, int blocks, cudaStream_t st)
{
    // one-time opt-in to the large dynamic shared memory, per device; callers may come from several host
    // threads (one per GPU in ssw_align_batch_multi), so the flags are atomics and a repeated call is harmless
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(score_kernel<K, TRUNC, REV, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCORE_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    score_kernel<K, TRUNC, REV, CHUNK><<<blocks, score_warps(K) * 32, SCORE_SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_k(const ScoreArgs& a, bool trunc, bool rev, int blocks, cudaStream_t st)
{
    // chunk mode (ScoreArgs::ck, long references) has its own instances, so that the common whole-pair
    // kernels carry none of its code
    if (a.ck.chunk_cols) {
        if (trunc) return rev ? launch_one<K, true, true, true>(a, blocks, st) : launch_one<K, true, false, true>(a, blocks, st);
        return rev ? launch_one<K, false, true, true>(a, blocks, st) : launch_one<K, false, false, true>(a, blocks, st);
    }
    if (trunc) return rev ? launch_one<K, true, true, false>(a, blocks, st) : launch_one<K, true, false, false>(a, blocks, st);
    return rev ? launch_one<K, false, true, false>(a, blocks, st) : launch_one<K, false, false, false>(a, blocks, st);
}

}  // namespace sswb
