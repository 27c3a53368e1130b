// ssw_band.cu -- banded affine DP + traceback -> CIGAR on the trimmed rectangle.
//
// Replaces banded_sw (reference ssw.c:548-735) as called from ssw_align (ssw.c:852-856).  The
// reference fills the band cell by cell in scalar int32 code; here one warp owns a pair and computes a
// whole band row at a time:
//   * cells are addressed by their band diagonal kk = j - i + w, so the vertical neighbour of slot kk is
//     slot kk+1 of the previous row and the diagonal neighbour is slot kk itself: the row buffers are
//     updated in place;
//   * the horizontal-gap chain F (a serial dependency along the row) is an exclusive max-plus prefix
//     scan over the lanes (valid because gap_open >= gap_extend on this path);
//   * one direction byte per cell (vertical source, horizontal source, H source) goes to a per-warp
//     scratch matrix; the band doubles while max < score1 exactly like ssw.c:631-632;
//   * lane 0 walks the matrix back from the bottom-right corner with the reference's state machine
//     (ssw.c:642-696), run-length encodes, and the warp writes the reversed ops to the output buffer.
// Tie rules, the "stop at read row 0" rule and the zeroed vertical neighbour of the last column in the
// first w+1 rows (ssw.c:595-596) are reproduced; see oracle/ssw_oracle.c:orc_band_cigar.
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

constexpr int NEG_INF = -(1 << 29);

// direction byte: bit0 vertical gap opened (code 3) / extended (2); bit1 horizontal gap opened (5) /
// extended (4); bits 2-3 H source: 0 diagonal, 1 vertical gap, 2 horizontal gap
__device__ __forceinline__ uint32_t cigar_pack(uint32_t len, int op) { return (len << 4) | (uint32_t)op; }

__device__ void band_pair(const BandArgs& a, const int pair, unsigned char* ws)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    PairRec* rec = a.b.rec + pair;
    const int refBeg = rec->ref_begin1, readBeg = rec->read_begin1;
    const int score = rec->score1;
    const int go = a.sc.go, ge = a.sc.ge;

    uint32_t* stage = reinterpret_cast<uint32_t*>(ws);                         // cigar ops, traceback order
    int* Hrow = reinterpret_cast<int*>(ws + (size_t)a.cigar_stage_cap * 4);
    // row buffers sized for the widest band that fits; direction matrix behind them
    int nOps = 0;
    int status = 0;

    if (refBeg < 0) {
        // score 0 in the byte flavour: the reference reads ref[-1] (undefined) on a 1x1 rectangle; the
        // traceback loop never runs and the CIGAR is "1M" (oracle/ssw_oracle.c documents the same rule).
        if (lane == 0) stage[0] = cigar_pack(1, 0);
        nOps = 1;
    } else {
        const int refLen = rec->ref_end1 - refBeg + 1;
        const int readLen = rec->read_end1 - readBeg + 1;
        const int8_t* ref = a.b.seqs + a.b.r_off[pair] + refBeg;
        const int8_t* read = a.b.seqs + a.b.q_off[pair] + readBeg;
        int bw = refLen - readLen; if (bw < 0) bw = -bw; bw += 1;
        int maxv = 0;
        int Wd = 0;
        unsigned char* dir = nullptr;
        for (;;) {
            Wd = 2 * bw + 1;
            const long long rowBufBytes = ((long long)(2 * bw + 4) * 8 + 15) & ~15LL;
            const long long need = rowBufBytes + (long long)Wd * readLen;
            if (need > a.dir_bytes) { status = PS_BAND_SCRATCH; break; }
            int* Erow = Hrow + (2 * bw + 4);
            dir = reinterpret_cast<unsigned char*>(Hrow) + rowBufBytes;
            for (int k = lane; k < 2 * bw + 4; k += 32) { Hrow[k] = 0; Erow[k] = 0; }
            __syncwarp();

            for (int i = 0; i < readLen; ++i) {
                const int beg = i - bw > 0 ? i - bw : 0;
                const int end = i + bw < refLen - 1 ? i + bw : refLen - 1;
                if (beg > end) break;                                   // the band has left the rectangle
                const int W = end - beg + 1;
                const int kk0 = beg - i + bw;                           // band diagonal of column beg
                int rd = read[i]; if ((unsigned)rd > 4u) rd = 4;
                // zeroed vertical neighbour of the last column (ssw.c:595-596): rows 1..w+1 whose band is
                // clipped by the reference end in this row and the previous one
                const bool quirk = i >= 1 && i - 1 - bw <= 0 && i - 1 + bw >= refLen - 1;
                unsigned char* drow = dir + (size_t)i * Wd;
                int carryV = NEG_INF;                                   // running max of A(k) + k*ge over earlier chunks
                int carryH = 0, carryF = 0;                             // H, F of the cell left of this chunk
                for (int t0 = 0; t0 < W; t0 += 32) {
                    const int t = t0 + lane;
                    const bool act = t < W;
                    const int kk = kk0 + t;
                    const int j = beg + t;
                    int Hd = 0, Hu = 0, Eu = 0, s = 0;
                    if (act) {
                        Hd = Hrow[kk]; Hu = Hrow[kk + 1]; Eu = Erow[kk + 1];
                        if (quirk && j == refLen - 1) { Hu = 0; Eu = 0; }
                        int rf = ref[j]; if ((unsigned)rf > 4u) rf = 4;
                        s = a.sc.mat[rf * 5 + rd];
                    }
                    __syncwarp();
                    const int eopen = Hu - go, eext = Eu - ge;
                    const int E = eopen > eext ? eopen : eext;
                    const int de = eopen > eext ? 1 : 0;
                    const int e1 = E > 0 ? E : 0;
                    const int dg = Hd + s;
                    const int A = e1 > dg ? e1 : dg;                    // H without the horizontal gap (>= 0)
                    // exclusive prefix max of V = A + t*ge
                    int V = act ? A + t * ge : NEG_INF;
                    int inc = V;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(FULL, inc, d);
                        if (lane >= d && o > inc) inc = o;
                    }
                    int exc = __shfl_up_sync(FULL, inc, 1);
                    if (lane == 0) exc = NEG_INF;
                    if (carryV > exc) exc = carryV;
                    // F(t) = max( -ge*(t+1), max_{k<t} (A(k) - go - (t-1-k)*ge) )
                    int F = -ge * (t + 1);
                    if (exc > NEG_INF) { const int f2 = exc - go - (t - 1) * ge; if (f2 > F) F = f2; }
                    const int H = A > F ? A : F;
                    // horizontal source: needs exact H, F of the left neighbour
                    int Hl = __shfl_up_sync(FULL, H, 1), Fl = __shfl_up_sync(FULL, F, 1);
                    if (lane == 0) { Hl = carryH; Fl = carryF; }
                    const int df = (Hl - go > Fl - ge) ? 1 : 0;
                    const int f1 = F > 0 ? F : 0;
                    const int gapbest = e1 > f1 ? e1 : f1;
                    int dh = 0;
                    if (gapbest > dg) dh = e1 > f1 ? 1 : 2;
                    if (act) {
                        Hrow[kk] = H; Erow[kk] = E;
                        drow[kk] = (unsigned char)(de | (df << 1) | (dh << 2));
                        if (H > maxv) maxv = H;
                    }
                    carryV = __shfl_sync(FULL, inc, 31) > carryV ? __shfl_sync(FULL, inc, 31) : carryV;
                    carryH = __shfl_sync(FULL, H, 31);
                    carryF = __shfl_sync(FULL, F, 31);
                    __syncwarp();
                }
            }
            maxv = __reduce_max_sync(FULL, maxv);
            bw *= 2;
            if (!(maxv < score && bw < 2 * readLen)) break;
        }
        bw /= 2;

        if (!status) {
            // traceback (lane 0): start in state H at the bottom-right corner, stop at read row 0
            if (lane == 0) {
                int i = readLen - 1, j = refLen - 1, state = 2;
                int op = 0, prevOp = 0, run = 0;                         // 0 M, 1 I, 2 D
                while (i > 0) {
                    const int beg = i - bw > 0 ? i - bw : 0;
                    const int end = i + bw < refLen - 1 ? i + bw : refLen - 1;
                    if (j < beg || j > end) { status = PS_TRACEBACK_ERR; break; }
                    const int d = dir[(size_t)i * Wd + (j - i + bw)];
                    int code;
                    if (state == 2) { const int dh = d >> 2; code = dh == 0 ? 1 : (dh == 1 ? 2 + (d & 1) : 4 + ((d >> 1) & 1)); }
                    else if (state == 0) code = 2 + (d & 1);
                    else code = 4 + ((d >> 1) & 1);
                    switch (code) {
                        case 1: --i; --j; state = 2; op = 0; break;
                        case 2: --i; state = 0; op = 1; break;
                        case 3: --i; state = 2; op = 1; break;
                        case 4: --j; state = 1; op = 2; break;
                        default: --j; state = 2; op = 2; break;
                    }
                    if (op == prevOp) ++run;
                    else {
                        if (nOps + 2 >= a.cigar_stage_cap) { status = PS_CIGAR_CAP; break; }
                        stage[nOps++] = cigar_pack((uint32_t)run, prevOp);
                        prevOp = op; run = 1;
                    }
                }
                if (!status) {
                    if (nOps + 2 >= a.cigar_stage_cap) status = PS_CIGAR_CAP;
                    else if (op == 0) stage[nOps++] = cigar_pack((uint32_t)run + 1, 0);     // ssw.c:697-704
                    else { stage[nOps++] = cigar_pack((uint32_t)run, op); stage[nOps++] = cigar_pack(1, 0); }
                }
            }
            nOps = __shfl_sync(FULL, nOps, 0);
            status = __shfl_sync(FULL, status, 0);
        }
    }

    long long off = 0;
    if (!status) {
        if (lane == 0) off = (long long)atomicAdd(a.cigar_used, (unsigned long long)nOps);
        off = __shfl_sync(FULL, off, 0);
        __syncwarp();
        if (off + nOps <= a.cigar_cap) {
            for (int k = lane; k < nOps; k += 32) a.cigar_buf[off + k] = stage[nOps - 1 - k];
        } else status = PS_CIGAR_CAP;
    }
    if (lane == 0) {
        rec->cigar_off = off;
        rec->cigar_len = status ? 0 : nOps;
        rec->status |= status;
    }
}

__global__ void __launch_bounds__(BAND_WARPS * 32) band_kernel(const BandArgs a)
{
    const int count = *a.wl.count;
    if (count <= 0) return;
    const int warp = threadIdx.x >> 5;
    const int base = a.wl.base ? *a.wl.base : 0;
    unsigned char* ws = a.scratch + (size_t)(blockIdx.x * BAND_WARPS + warp) * a.scratch_stride;
    for (;;) {
        int idx = 0;
        if (lane_id() == 0) idx = atomicAdd(a.wl.cursor, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        band_pair(a, a.wl.idx[base + idx], ws);
        __syncwarp();
    }
}

cudaError_t launch_band(const BandArgs& a, int blocks, cudaStream_t st)
{
    band_kernel<<<blocks, BAND_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sswb
