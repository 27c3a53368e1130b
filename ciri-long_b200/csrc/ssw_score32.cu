// ssw_score32.cu -- 32-bit score passes for the pairs the packed 16-bit kernels hand over.
//
// The packed kernels of ssw_score_impl.cuh keep scores in signed 16-bit halves and leave a pair alone as soon
// as its score comes near the range where the reference's own arithmetic starts to matter: the word
// flavour saturates at 32767 (`_mm_adds_epi16`, ssw.c:442), and the truncated-F gate of the packed
// kernel needs scores below 16000.  Such pairs (e.g. >3.3 kb near-perfect matches at match = 10) are
// re-done here with one strip per lane in int32 registers (32-bit DPX: VIADDMNMX / VIMNMX3) and the saturating add
// made explicit:
//     H = max(0, min(Hdiag + s, 32767), E, F)
// Same wavefront, same strip layouts (right-aligned strips for GOTOH, segment-aligned strips with
// arbitrary live counts for TRUNC), same outputs; throughput is secondary (a handful of pairs).
#include "ssw_common.cuh"
#include "ssw_kernels.h"
#include "ssw_second_best.cuh"

namespace sswb {

constexpr int K32 = 16;          // rows per strip
constexpr int V32 = 32;          // strips per tile (one per lane)

template <bool TRUNC, bool REV>
__device__ void score32_pair(const ScoreArgs& a, const int pair, unsigned char* ws, const int* matS)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    PairRec* rec = a.b.rec + pair;
    int m, n, terminate = 0;
    const int8_t* qb;
    const int8_t* rb;
    int qs, rs;
    if (!REV) {
        m = a.b.q_len[pair]; n = a.b.r_len[pair];
        qb = a.b.seqs + a.b.q_off[pair]; rb = a.b.seqs + a.b.r_off[pair];
        qs = 1; rs = 1;
    } else {
        m = rec->read_end1 + 1; n = rec->ref_end1 + 1;
        qb = a.b.seqs + a.b.q_off[pair] + rec->read_end1;
        rb = a.b.seqs + a.b.r_off[pair] + rec->ref_end1;
        qs = -1; rs = -1;
        terminate = rec->score1;
    }
    const int go = a.sc.go, ge = a.sc.ge;
    unsigned* colbuf = reinterpret_cast<unsigned*>(ws + a.off_col);
    uint2* bnd = reinterpret_cast<uint2*>(ws + a.off_bnd);

    int T, Vtot, dead = 0, segLen = 0, G = 1, base = 0, extra = 0;
    if (!TRUNC) {
        const int rpt = V32 * K32;
        T = (m + rpt - 1) / rpt; dead = T * rpt - m; Vtot = T * V32;
    } else {
        segLen = (m + 7) / 8;
        G = (segLen + K32 - 1) / K32;
        base = segLen / G; extra = segLen - base * G;
        Vtot = 8 * G; T = (Vtot + V32 - 1) / V32;
    }

    int candM = 0, candCol = -1, candRow = 0, termCol = -1, overCol = 0x7fffffff;
    bool exactMode = false;
    for (int attempt = 0; attempt < (REV ? 2 : 1); ++attempt) {
        candM = 0; candCol = -1; candRow = 0; termCol = -1; overCol = 0x7fffffff;
        for (int p = 0; p < T; ++p) {
            const bool lastTile = p == T - 1;
            const int v = p * V32 + lane;
            int first, live; bool valid, segStart = false;
            if (!TRUNC) { first = v * K32 - dead; live = K32; valid = true; }
            else {
                valid = v < Vtot;
                const int l = v / G, g = v - l * G;
                live = valid ? base + (g < extra ? 1 : 0) : 0;
                first = l * segLen + g * base + (g < extra ? g : extra);
                segStart = valid && g == 0 && l >= 1;
            }
            // (the first row that holds a column's maximum is tracked with the maximum itself, so the best cell needs no
            // snapshot of the strip's column: ssw.c:299-308 asks for the smallest row with H == max in the best column)
            int qc[K32], E[K32], Hd[K32];
#pragma unroll
            for (int i = 0; i < K32; ++i) {
                const int r = first + i;
                int c = 4;
                if (r >= 0 && r < m && i < live) { c = qb[(long long)r * qs]; if ((unsigned)c > 4u) c = 4; }
                qc[i] = c; E[i] = 0; Hd[i] = 0;
            }
            const int wv = (TRUNC && lastTile) ? Vtot - 1 - p * V32 : V32 - 1;
            int Hout = 0, Fout = 0, R = 0, diagIn = 0, best = 0, bcol = -1, brow = 0, termflag = 0;
            const int steps = n + wv;
            for (int s = 0; s < steps; ++s) {
                int rH = __shfl_up_sync(FULL, Hout, 1), rF = __shfl_up_sync(FULL, Fout, 1), rR = __shfl_up_sync(FULL, R, 1);
                if (lane == 0) {
                    rH = 0; rF = 0; rR = 0;
                    if (p > 0 && s < n) {
                        const uint2 bv = bnd[s];
                        rH = (int)(bv.x & 0xffffu); rF = (int)(short)(bv.x >> 16); rR = (int)bv.y;
                    }
                }
                const int c = s - lane;
                const bool colOk = (unsigned)c < (unsigned)n;
                int rc = 4;
                if (colOk) { rc = rb[(long long)c * rs]; if ((unsigned)rc > 4u) rc = 4; }
                const int* mrow = matS + rc * 5;                            // (N row / column of the matrix are zero: scoring_supported)
                int diag = diagIn; diagIn = rH;
                int F = rF, mx = 0, mrowIdx = 0, Hk = rH, Fk = rF;
#pragma unroll
                for (int i = 0; i < K32; ++i) {
                    if (i < live) {
                        const int x = __viaddmin_s32(diag, mrow[qc[i]], 32767);          // _mm_adds_epi16 (ssw.c:442)
                        const int h = __vimax3_s32_relu(x, E[i], F);
                        int u;
                        if (TRUNC && i == 0 && segStart) {                 // cut vertical-gap chain (ssw.c:467-478)
                            const int h0 = __vimax_s32_relu(x, E[i]);
                            u = h0 - go;
                            F = u;
                        } else {
                            u = h - go;
                            F = __viaddmax_s32(F, -ge, u);
                        }
                        E[i] = __viaddmax_s32(E[i], -ge, u);
                        diag = Hd[i]; Hd[i] = h;
                        if (h > mx) { mx = h; mrowIdx = i; }
                        if (i == live - 1) { Hk = h; Fk = F; }
                    }
                }
                Hout = Hk; Fout = Fk;
                if (!colOk || !valid) mx = 0;
                R = rR > mx ? rR : mx;
                if (REV && !exactMode && mx > terminate) { overCol = c < overCol ? c : overCol; mx = 0; }
                if (mx > best) { best = mx; bcol = c; brow = first + mrowIdx; }
                if (lane == wv && colOk) {
                    if (!lastTile) bnd[c] = make_uint2((unsigned)(Hout & 0xffff) | ((unsigned)(Fout & 0xffff) << 16), (unsigned)R);
                    else if (!REV) colbuf[c] = (unsigned)R | ((unsigned)Hout << 16);
                    else if (!termflag && !exactMode && R == terminate) { termflag = 1; termCol = c; }
                }
                if (REV && lastTile && (s & 7) == 7 && __any_sync(FULL, termflag)) break;
            }
            // tile epilogue (max, first column, first row)
            const int M = __reduce_max_sync(FULL, best);
            if (M > 0) {
                const int col = __reduce_min_sync(FULL, best == M ? bcol : 0x7fffffff);
                const int owner = __reduce_min_sync(FULL, (best == M && bcol == col) ? lane : 1000);
                int row = brow > m - 1 ? m - 1 : brow;
                row = __shfl_sync(FULL, row, owner);
                if (M > candM || (M == candM && col < candCol)) { candM = M; candCol = col; candRow = row; }
            }
            if (lastTile) termCol = __shfl_sync(FULL, termCol, wv);
            __syncwarp();
        }
        if (REV && !exactMode) {
            overCol = __reduce_min_sync(FULL, overCol);
            if (overCol != 0x7fffffff && (termCol < 0 || overCol <= termCol)) {
                exactMode = true;
                if (termCol >= 0) n = termCol + 1;
                continue;
            }
        }
        break;
    }

    if (!REV) {
        const bool over8 = candM + a.sc.bias >= 255;
        const int word = TRUNC ? 1 : (over8 ? 1 : 0);
        const int endRef = candM > 0 ? candCol : (word ? 0 : -1);
        const int endRead = candM > 0 ? candRow : 0;
        int score2 = 0, ref2 = -1;
        const int maskLen = a.b.mask_len[pair];
        if (maskLen >= 15) {
            ref2 = 0;
            const int L = word ? 8 : 16;
            const int P = TRUNC ? 0 : ((m + L - 1) / L) * L - m;
            second_best(colbuf, n, P, 0, word, endRef, maskLen, go, ge, lane, score2, ref2);
        }
        if (lane == 0) {
            rec->score1 = candM; rec->score2 = score2;
            rec->ref_begin1 = -1; rec->ref_end1 = endRef;
            rec->read_begin1 = -1; rec->read_end1 = endRead;
            rec->ref_end2 = ref2; rec->cigar_len = 0; rec->cigar_off = 0;
            rec->word = word;
            rec->status = PS_WIDE32;
        }
    } else if (lane == 0) {
        const int word = rec->word;
        rec->ref_begin1 = candM > 0 ? rec->ref_end1 - candCol : (word ? 0 : -1);
        rec->read_begin1 = rec->read_end1 - (candM > 0 ? candRow : 0);
    }
}

template <bool TRUNC, bool REV>
__global__ void __launch_bounds__(SCORE32_WARPS * 32) score32_kernel(const ScoreArgs a)
{
    __shared__ int matS[25];
    const int count = *a.wl.count;
    if (count <= 0) return;
    if (threadIdx.x < 25) matS[threadIdx.x] = a.sc.mat[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int base = a.wl.base ? *a.wl.base : 0;
    unsigned char* ws = a.scratch + (size_t)(blockIdx.x * SCORE32_WARPS + warp) * a.scratch_stride;
    for (;;) {
        int idx = 0;
        if (lane_id() == 0) idx = atomicAdd(a.wl.cursor, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        score32_pair<TRUNC, REV>(a, a.wl.idx[base + idx], ws, matS);
        __syncwarp();
    }
}

cudaError_t launch_score32(bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st)
{
    if (trunc) { if (rev) score32_kernel<true, true><<<blocks, SCORE32_WARPS * 32, 0, st>>>(a); else score32_kernel<true, false><<<blocks, SCORE32_WARPS * 32, 0, st>>>(a); }
    else { if (rev) score32_kernel<false, true><<<blocks, SCORE32_WARPS * 32, 0, st>>>(a); else score32_kernel<false, false><<<blocks, SCORE32_WARPS * 32, 0, st>>>(a); }
    return cudaGetLastError();
}

}  // namespace sswb
