// ssw_score.cu -- score passes of the striped Smith-Waterman hot path for sm_100a.
//
// Replaces sw_sse2_byte / sw_sse2_word (reference ssw.c:123-345, 371-546) for the forward pass and
// for the reverse pass of ssw_align (ssw.c:836-849).  Not a port of the SSE2 code:
//
//   * inter-task parallelism, one warp per (query, reference) pair;
//   * the query is cut into 64 "virtual strips" of K consecutive rows: lane t owns strip t in the low
//     16-bit half and strip t+32 in the high half of every register, so each DPX instruction
//     (VIADDMNMX.S16x2 / VIMNMX.S16x2) updates two cells;
//   * a systolic wavefront over the reference: at step s strip v computes column s-v; H, F and the
//     running column maximum leave a strip through one rotate-shuffle per value and step;
//   * substitution scores come from a lane-replicated (bank == lane, conflict free) shared-memory
//     table indexed by (ref base pair, query base pair): one LDS per two cells, off the DPX pipe;
//   * queries longer than 64*K rows are processed in row tiles that hand H/F/colmax to the next tile
//     through a per-warp boundary array.
//
// What is computed is the *semantics* of the reference, verified against oracle/ssw_oracle.c:
//   GOTOH  plain affine-gap recurrences on the real query rows (equal to the reference's byte flavour,
//          and to its word flavour when gap_open > gap_extend), the flavour (8/16 bit) being decided
//          after the pass from max+bias >= 255 exactly like ssw.c:285,317,806;
//   TRUNC  the word flavour when gap_open == gap_extend: the reference's lazy-F loop stops after one
//          step (ssw.c:467-478), so the vertical-gap chain is cut at every segment boundary
//          (row % ceil(m/8) == 0) and only the boundary row's H sees the incoming F.
// The reference pads the query to a multiple of 16 (8) rows with zero-scoring rows; those rows only
// influence maxColumn[] (second-best score).  Their contribution is added in closed form from the
// last real row (see second_best()), so the DP itself runs on real rows only.
#include <stdio.h>
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

// Second-best score outside the mask window around the best end column (ssw.c:325-340 / 528-541).
// colbuf[c] = colmax over the real query rows | H(last real row, c) << 16.  The reference's maxColumn[]
// also covers P zero-scoring pad rows (P = L*ceil(m/L) - m, L = 16 byte / 8 word flavour).  A pad cell
// is reached from the last real row by free diagonal steps plus at most one horizontal gap, so
//   padmax(c) = max( max_{1<=d<=P} Hlast(c-d),  G(c) ),   G(c) = max(G(c-1) - ge, Hlast(c-P-1) - go, 0)
// (vertical gaps are dominated inside the same column).  Lanes take contiguous column chunks; the
// decaying chain G is stitched across chunks with one pass over the per-lane carries.
__device__ __noinline__ void second_best(const unsigned* colbuf, int n, int m, int word, int endRef, int maskLen,
                                         int go, int ge, int lane, int& score2, int& ref2)
{
    const unsigned FULL = 0xffffffffu;
    const int L = word ? 8 : 16;
    const int P = ((m + L - 1) / L) * L - m;
    const int e1 = endRef - maskLen > 0 ? endRef - maskLen : 0;              // left region  [0, e1)
    int e2 = endRef + maskLen > n ? n : endRef + maskLen;                    // right region [e2, n)
    if (!word) e2 += 1;                                                      // ssw.c:334 vs ssw.c:536
    const int chunk = (n + 31) / 32;
    const int c0 = lane * chunk < n ? lane * chunk : n;
    const int c1 = c0 + chunk < n ? c0 + chunk : n;

    int gin = 0;
    if (P > 0) {
        int carry = 0;
        for (int c = c0; c < c1; ++c) {
            const int src = c - P - 1 >= 0 ? (int)(colbuf[c - P - 1] >> 16) - go : -1;
            carry = carry - ge > src ? carry - ge : src;
            if (carry < 0) carry = 0;
        }
        int run = 0;
        for (int l = 0; l < 32; ++l) {
            const int cl = __shfl_sync(FULL, carry, l);
            const int len = __shfl_sync(FULL, c1 - c0, l);
            if (lane == l) gin = run;
            run = run - ge * len > cl ? run - ge * len : cl;
            if (run < 0) run = 0;
        }
    }
    int bv = 0, bi = 0x7fffffff, G = gin;
    for (int c = c0; c < c1; ++c) {
        int mc = (int)(colbuf[c] & 0xffffu);
        if (P > 0) {
            const int src = c - P - 1 >= 0 ? (int)(colbuf[c - P - 1] >> 16) - go : -1;
            G = G - ge > src ? G - ge : src;
            if (G < 0) G = 0;
            int D = G;
            const int dmax = P < c ? P : c;
            for (int d = 1; d <= dmax; ++d) { const int v = (int)(colbuf[c - d] >> 16); D = v > D ? v : D; }
            mc = D > mc ? D : mc;
        }
        if ((c < e1 || c >= e2) && mc > bv) { bv = mc; bi = c; }
    }
    const int M = __reduce_max_sync(FULL, bv);
    const int idx = __reduce_min_sync(FULL, (bv == M && M > 0) ? bi : 0x7fffffff);
    score2 = M;
    ref2 = M > 0 ? idx : 0;
}

template <int K, bool TRUNC, bool REV>
__device__ __forceinline__ void score_pair(const ScoreArgs& a, const int pair, const unsigned* lut, unsigned char* ws)
{
    const unsigned FULL = 0xffffffffu;
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

    // per-warp scratch: [ref-pair codes | column records | tile boundary | best-column snapshots]
    unsigned char* rpbuf = ws;
    unsigned* colbuf = reinterpret_cast<unsigned*>(ws + a.off_col);
    uint2* bnd = reinterpret_cast<uint2*>(ws + a.off_bnd);
    unsigned* snap = reinterpret_cast<unsigned*>(ws + a.off_snap);

    // rpbuf[x], x = c + 32: code(c) * 5 + code(c - 32); columns outside [0, n) read as Z (score 0)
    for (int x = lane; x < n + 96; x += 32) {
        const int c = x - 32, c2 = x - 64;
        int ca = 4, cb = 4;
        if (c >= 0 && c < n) { ca = rb[(long long)c * rs]; if ((unsigned)ca > 4u) ca = 4; }
        if (c2 >= 0 && c2 < n) { cb = rb[(long long)c2 * rs]; if ((unsigned)cb > 4u) cb = 4; }
        rpbuf[x] = (unsigned char)(ca * 5 + cb);
    }
    __syncwarp();

    const int rpt = VSTRIPS * K;                      // rows per tile
    const int T = (m + rpt - 1) / rpt;
    const int dead = T * rpt - m;                      // zero rows in front of row 0 (right-aligned strips)
    const int segLen = (m + 7) / 8;
    const unsigned mgo = pack2(-a.sc.go, -a.sc.go), mge = pack2(-a.sc.ge, -a.sc.ge);
    const int src = (lane + 31) & 31;
    const unsigned fix = lane == 0 ? 0x1044u : 0x3210u;   // lane 0: high half <- lane 31's low half, low half <- 0
    const char* lutbase = reinterpret_cast<const char*>(lut) + lane * 4;

    int candM = 0, candCol = -1, candRow = 0;
    int termCol = -1, overCol = 0x7fffffff;
    const unsigned mtermP = pack2(-terminate, -terminate);

    for (int p = 0; p < T; ++p) {
        const bool lastTile = (p == T - 1);
        const int rowbase = p * rpt - dead;
        const int r0lo = rowbase + lane * K, r0hi = rowbase + (lane + 32) * K;
        unsigned qoff[K], E[K], Hd[K];
        unsigned g[TRUNC ? K : 1], gF[TRUNC ? K : 1];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int rl = r0lo + i, rh = r0hi + i;
            int cl = 4, ch = 4;
            if (rl >= 0) { cl = qb[(long long)rl * qs]; if ((unsigned)cl > 4u) cl = 4; }
            if (rh >= 0) { ch = qb[(long long)rh * qs]; if ((unsigned)ch > 4u) ch = 4; }
            qoff[i] = (unsigned)(cl * 5 + ch) * 128u;
            E[i] = 0; Hd[i] = 0;
            if (TRUNC) {
                const bool bl = rl > 0 && (rl % segLen) == 0, bh = rh > 0 && (rh % segLen) == 0;
                g[i] = pack2(bl ? TRUNC_GATE : 0, bh ? TRUNC_GATE : 0);
                gF[i] = pack2(bl ? TRUNC_GATE : -a.sc.ge, bh ? TRUNC_GATE : -a.sc.ge);
            }
        }
        unsigned Hout = 0, Fout = 0, R = 0, diagIn = 0, best = 0;
        int bcolLo = -1, bcolHi = -1;
        int cLo = -lane, cHi = -lane - 32;
        int termflag = 0;
        const int steps = n + 63;
        unsigned rpCur = rpbuf[32 - lane];
        unsigned rpNxt = rpbuf[33 - lane];

        for (int s = 0; s < steps; ++s) {
            // hand-off from the previous strip (computed one step ago, same column as ours now)
            unsigned rH = __byte_perm(__shfl_sync(FULL, Hout, src), 0u, fix);
            unsigned rF = __byte_perm(__shfl_sync(FULL, Fout, src), 0u, fix);
            unsigned rR = __byte_perm(__shfl_sync(FULL, R, src), 0u, fix);
            if (p > 0 && lane == 0 && s < n) {           // strip 0 continues below the previous tile
                const uint2 bv = bnd[s];
                rH |= bv.x & 0xffffu; rF |= bv.x >> 16; rR |= bv.y & 0xffffu;
            }
            unsigned diag = diagIn;
            diagIn = rH;
            unsigned F = rF;
            const char* sb = lutbase + rpCur * RP_STRIDE;
            rpCur = rpNxt;
            { int x = s + 34 - lane; x = x < n + 95 ? x : n + 95; rpNxt = rpbuf[x]; }

#ifdef SSW_DEBUG
            if (rpCur > 24u) { printf("BAD rp %u pair %d lane %d s %d n %d m %d p %d K %d rev %d\n", rpCur, pair, lane, s, n, m, p, K, (int)REV); rpCur = 24; }
#endif
            unsigned mx = 0, hprev = 0;
#pragma unroll
            for (int i = 0; i < K; ++i) {
#ifdef SSW_DEBUG
                if (qoff[i] > 3072u) { printf("BAD qoff %u pair %d lane %d i %d\n", qoff[i], pair, lane, i); }
#endif
                const unsigned sc = *reinterpret_cast<const unsigned*>(sb + qoff[i]);
                const unsigned x = addmax(diag, sc, E[i]);
                const unsigned h = max_relu(x, F);
                unsigned u;
                if (TRUNC) {
                    const unsigned h0 = addmax_relu(F, g[i], x);
                    u = addmax(h0, mgo, S16X2_MIN);
                    F = addmax(F, gF[i], u);
                } else {
                    u = addmax(h, mgo, S16X2_MIN);
                    F = addmax(F, mge, u);
                }
                E[i] = addmax(E[i], mge, u);
                diag = Hd[i];
                Hd[i] = h;
                if (i & 1) mx = max3(mx, hprev, h);
                hprev = h;
            }
            if (K & 1) mx = max_relu(mx, hprev);
            Hout = Hd[K - 1];
            Fout = F;

            const unsigned vm = ((unsigned)cLo < (unsigned)n ? 0xffffu : 0u) | ((unsigned)cHi < (unsigned)n ? 0xffff0000u : 0u);
            unsigned mxv = mx & vm;
            R = max_relu(rR, mxv);
            if (REV) {
                // The reference stops at the first column whose maximum equals score1 (ssw.c:296,499), so
                // cells above score1 only count if they occur before that column.  Strips ahead of the
                // stop column keep running here: values above score1 are kept out of the best-cell
                // tracking and only their first column is remembered (checked after the pass).
                // (the DPX result must be consumed: ptxas 12.9 mis-allocates the destination of a VIMNMX whose
                //  value is dead and only the predicates are used)
                const unsigned ov = addmax_relu(mxv, mtermP, 0u);              // max(mxv - score1, 0) per half
                if (ov) {
                    if (ov & 0xffffu) { overCol = cLo < overCol ? cLo : overCol; mxv &= 0xffff0000u; }
                    if (ov >> 16) { overCol = cHi < overCol ? cHi : overCol; mxv &= 0x0000ffffu; }
                }
            }
            bool pHi, pLo;
            const unsigned nb = __vibmax_s16x2(best, mxv, &pHi, &pLo);   // pred = (best >= mxv)
            if (!(pHi && pLo)) {
                if (!pLo) {
                    bcolLo = cLo;
#pragma unroll
                    for (int i = 0; i < K; ++i) snap[i * 32 + lane] = Hd[i];
                }
                if (!pHi) {
                    bcolHi = cHi;
#pragma unroll
                    for (int i = 0; i < K; ++i) snap[(K + i) * 32 + lane] = Hd[i];
                }
                best = nb;
            }
            if (lane == 31 && (unsigned)cHi < (unsigned)n) {
                if (!lastTile) {
                    bnd[cHi] = make_uint2(__byte_perm(Hout, Fout, 0x7632u), R >> 16);
                } else if (!REV) {
                    colbuf[cHi] = __byte_perm(R, Hout, 0x7632u);           // colmax | H(last row) << 16
                } else if (!termflag && (int)(R >> 16) == terminate) {
                    termflag = 1; termCol = cHi;
                }
            }
            ++cLo; ++cHi;
            if (REV && lastTile && (s & 7) == 7) {
                if (__any_sync(FULL, termflag)) break;
            }
        }

        // ---- tile epilogue: best cell of this tile in reference order (max, first column, first row)
        const int vlo = lo16(best), vhi = hi16(best);
        const int M = __reduce_max_sync(FULL, vlo > vhi ? vlo : vhi);
        if (M > 0) {
            const int clo = vlo == M ? bcolLo : 0x7fffffff, chi = vhi == M ? bcolHi : 0x7fffffff;
            const int col = __reduce_min_sync(FULL, clo < chi ? clo : chi);
            const int st = (vlo == M && bcolLo == col) ? lane : ((vhi == M && bcolHi == col) ? lane + 32 : 1000);
            const int strip = __reduce_min_sync(FULL, st);
            const int owner = strip & 31, half = strip >> 5;
            int row = 0;
            if (lane == owner) {
                for (int i = K - 1; i >= 0; --i) {
                    const unsigned v = snap[(half * K + i) * 32 + lane];
                    if ((half ? hi16(v) : lo16(v)) == M) row = rowbase + strip * K + i;
                }
            }
            row = __shfl_sync(FULL, row, owner);
            if (M > candM || (M == candM && col < candCol)) { candM = M; candCol = col; candRow = row; }
        }
        __syncwarp();      // boundary array / snapshots written by this tile are read by the next one
    }

    if (!REV) {
        const bool over8 = candM + a.sc.bias >= 255;               // ssw.c:285,317
        int word = TRUNC ? 1 : (over8 ? 1 : 0);
        int status = 0;
        if (!TRUNC && a.rerun) {
            // second look at a pair whose truncated pass stayed below the 8-bit limit: the byte flavour is
            // authoritative unless it overflows, in which case the word result already stored stands.
            if (over8) { if (lane == 0) rec->status &= ~PS_NEED_GOTOH; return; }
        } else if (!TRUNC && over8 && a.sc.go == a.sc.ge) {
            status |= PS_PUNT;          // host routing guarantees this cannot happen; never guess
        }
        if (TRUNC && !over8) status |= PS_NEED_GOTOH;
        if (candM >= (TRUNC ? TRUNC_SCORE_LIMIT : S16_SCORE_LIMIT)) status |= PS_PUNT;

        const int endRef = candM > 0 ? candCol : (word ? 0 : -1);  // ssw.c:145 vs ssw.c:388
        const int endRead = candM > 0 ? candRow : 0;
        int score2 = 0, ref2 = -1;
        const int maskLen = a.b.mask_len[pair];
        if (maskLen >= 15) {                                        // ssw.c:826-832
            ref2 = 0;
            second_best(colbuf, n, m, word, endRef, maskLen, a.sc.go, a.sc.ge, lane, score2, ref2);
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
        }
    } else {
        int status = 0;
        // a cell above score1 before the stop column (or with no stop column at all): the truncated-F
        // flavour scored the prefix higher than the forward pass did; the exact kernel decides
        termCol = __shfl_sync(FULL, termCol, 31);
        overCol = __reduce_min_sync(FULL, overCol);
        if (overCol != 0x7fffffff && (termCol < 0 || overCol <= termCol)) status |= PS_PUNT;
        if (lane == 0) {
            const int word = rec->word;
            rec->ref_begin1 = candM > 0 ? rec->ref_end1 - candCol : (word ? 0 : -1);
            rec->read_begin1 = rec->read_end1 - (candM > 0 ? candRow : 0);
            rec->status |= status;
        }
    }
}

template <int K, bool TRUNC, bool REV>
__global__ void __launch_bounds__(SCORE_THREADS, 1) score_kernel(const ScoreArgs a)
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
    unsigned char* ws = a.scratch + (size_t)(blockIdx.x * SCORE_WARPS + warp) * a.scratch_stride;
    for (;;) {
        int idx = 0;
        if (lane_id() == 0) idx = atomicAdd(a.wl.cursor, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        score_pair<K, TRUNC, REV>(a, a.wl.idx[base + idx], lut, ws);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// host-side launch table

template <int K, bool TRUNC, bool REV>
static cudaError_t launch_one(const ScoreArgs& a, int blocks, cudaStream_t st)
{
    static bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(score_kernel<K, TRUNC, REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_BYTES);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    score_kernel<K, TRUNC, REV><<<blocks, SCORE_THREADS, LUT_BYTES, st>>>(a);
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_k(const ScoreArgs& a, bool trunc, bool rev, int blocks, cudaStream_t st)
{
    if (trunc) return rev ? launch_one<K, true, true>(a, blocks, st) : launch_one<K, true, false>(a, blocks, st);
    return rev ? launch_one<K, false, true>(a, blocks, st) : launch_one<K, false, false>(a, blocks, st);
}

cudaError_t launch_score(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st)
{
    switch (K) {
        case 1: return launch_k<1>(a, trunc, rev, blocks, st);
        case 2: return launch_k<2>(a, trunc, rev, blocks, st);
        case 3: return launch_k<3>(a, trunc, rev, blocks, st);
        case 4: return launch_k<4>(a, trunc, rev, blocks, st);
        case 5: return launch_k<5>(a, trunc, rev, blocks, st);
        case 6: return launch_k<6>(a, trunc, rev, blocks, st);
        case 7: return launch_k<7>(a, trunc, rev, blocks, st);
        case 8: return launch_k<8>(a, trunc, rev, blocks, st);
        case 9: return launch_k<9>(a, trunc, rev, blocks, st);
        case 10: return launch_k<10>(a, trunc, rev, blocks, st);
        case 11: return launch_k<11>(a, trunc, rev, blocks, st);
        case 12: return launch_k<12>(a, trunc, rev, blocks, st);
        case 13: return launch_k<13>(a, trunc, rev, blocks, st);
        case 14: return launch_k<14>(a, trunc, rev, blocks, st);
        case 15: return launch_k<15>(a, trunc, rev, blocks, st);
        case 16: return launch_k<16>(a, trunc, rev, blocks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sswb
