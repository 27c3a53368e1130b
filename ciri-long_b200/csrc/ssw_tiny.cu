// ssw_tiny.cu -- score passes for tiny pairs, ONE PAIR PER THREAD: forward pass, second-best scan and the reverse
// pass of ssw_align (reference ssw.c:123-345, 325-340, 836-849) fused in one kernel.
//
// collapse.curate_junction (CIRI_long/collapse.py:165-172) aligns a ~50-nt junction consensus against thousands of
// 20-nt genomic junctions per cluster; a warp-wide systolic pass (ssw_score_impl.cuh) spends most of its 64 strips
// and its pipeline fill on such a pair.  A pair qualifies when both sequences have at most TINY_LEN bases and
// min(m, n) * (largest score) + bias < 255: its score cannot leave the 8-bit range, so the reference's byte
// flavour is the one that answers (ssw.c:805-809), and that flavour equals the plain affine-gap recurrences on
// the real query rows -- no truncated-F case, no re-run, no saturation.
//   * H|E of the previous column live in shared memory, indexed [row][thread]: conflict free;
//   * column maximum and the first row that holds it are tracked per column, which gives the reference's best
//     cell (largest score, first column, smallest row; ssw.c:283-308) without a snapshot;
//   * the second-best scan (ssw.c:325-340) with the byte flavour's 16-row padding in closed form, like
//     ssw_second_best.cuh, from per-column (maximum, H of the last row) records in shared memory;
//   * the reverse pass runs on the reversed prefixes and stops at the first column whose maximum equals the
//     score (ssw.c:296).
// The CIGAR comes from the ordinary CIGAR stage (ssw_tband.cu is itself one pair per lane).
#include <atomic>
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

constexpr int TINY_THREADS = 128;

struct TinyBest { int score, col, row; };

// One score pass of a thread.  q/r: first base of query / reference in walking order, qs/rs: +1 or -1.
// HE: this thread's column state, HE[i * TINY_THREADS] = H | E << 8 (both in 0..254: the pair cannot leave the 8-bit
// range).  colrec (forward only): colrec[c * TINY_THREADS] = column maximum | H(last row) << 8.  terminate > 0: stop at the first column whose
// maximum equals it.
template <bool REV>
__device__ __forceinline__ TinyBest tiny_pass(const int8_t* q, const int qs, const int m, const int8_t* r, const int rs, const int n,
                                              const int* matS, const int go, const int ge, unsigned short* HE, unsigned short* colrec,
                                              const unsigned char* qcode, const int terminate)
{
    for (int i = 0; i < m; ++i) HE[i * TINY_THREADS] = 0;
    TinyBest best = {0, -1, 0};
    for (int j = 0; j < n; ++j) {
        unsigned rc = (unsigned char)r[(long long)j * rs]; if (rc > 4u) rc = 4;
        const int* mrow = matS + rc * 5;
        int F = 0, diag = 0, h = 0, cm = 0, cmRow = 0;
        for (int i = 0; i < m; ++i) {
            const unsigned he = HE[i * TINY_THREADS];
            const int hl = (int)(he & 0xffu), e = (int)(he >> 8);           // H(i, j-1), E(i, j) computed after column j-1
            h = diag + mrow[qcode[i * TINY_THREADS]];
            h = h > e ? h : e;
            h = h > F ? h : F;
            h = h > 0 ? h : 0;
            const int t = h - go;
            int en = e - ge; en = en > t ? en : t; en = en > 0 ? en : 0;
            F = F - ge; F = F > t ? F : t; F = F > 0 ? F : 0;
            HE[i * TINY_THREADS] = (unsigned short)((unsigned)h | ((unsigned)en << 8));
            if (h > cm) { cm = h; cmRow = i; }
            diag = hl;
        }
        if (!REV) colrec[j * TINY_THREADS] = (unsigned short)((unsigned)cm | ((unsigned)h << 8));
        if (cm > best.score) { best.score = cm; best.col = j; best.row = cmRow; }
        if (REV && cm == terminate) break;
    }
    return best;
}

__global__ void __launch_bounds__(TINY_THREADS) tiny_kernel(const TinyArgs a)
{
    extern __shared__ unsigned tiny_smem[];
    __shared__ int matS[25];
    const int count = *a.wl.count;
    if (count <= 0) return;
    if (threadIdx.x < 25) matS[threadIdx.x] = a.sc.mat[threadIdx.x];
    __syncthreads();
    unsigned short* HE = reinterpret_cast<unsigned short*>(tiny_smem) + threadIdx.x;
    unsigned short* colrec = HE + TINY_LEN * TINY_THREADS;
    unsigned char* qcode = reinterpret_cast<unsigned char*>(colrec - threadIdx.x + TINY_LEN * TINY_THREADS) + threadIdx.x;
    const int base = a.wl.base ? *a.wl.base : 0;
    const int go = a.sc.go, ge = a.sc.ge;
    for (int k = blockIdx.x * TINY_THREADS + threadIdx.x; k < count; k += gridDim.x * TINY_THREADS) {
        const int pair = a.wl.idx[base + k];
        PairRec* rec = a.b.rec + pair;
        const int m = a.b.q_len[pair], n = a.b.r_len[pair];
        const int8_t* qb = a.b.seqs + a.b.q_off[pair];
        const int8_t* rb = a.b.seqs + a.b.r_off[pair];
        for (int i = 0; i < m; ++i) { unsigned c = (unsigned char)qb[i]; qcode[i * TINY_THREADS] = (unsigned char)(c > 4u ? 4u : c); }
        const TinyBest f = tiny_pass<false>(qb, 1, m, rb, 1, n, matS, go, ge, HE, colrec, qcode, 0);
        const int endRef = f.score > 0 ? f.col : -1;                     // byte flavour: end_ref starts at -1 (ssw.c:145)
        const int endRead = f.score > 0 ? (f.row > m - 1 ? m - 1 : f.row) : 0;
        // second best outside the mask window (ssw.c:325-340), 16-row padding in closed form (ssw_second_best.cuh)
        int score2 = 0, ref2 = -1;
        const int maskLen = a.b.mask_len[pair];
        if (maskLen >= 15) {
            ref2 = 0;
            const int P = ((m + 15) / 16) * 16 - m;
            const int e1 = endRef - maskLen > 0 ? endRef - maskLen : 0;
            int e2 = endRef + maskLen > n ? n : endRef + maskLen;
            e2 += 1;
            int bv = 0, bi = 0, G = 0;
            for (int c = 0; c < n; ++c) {
                int mc = (int)(colrec[c * TINY_THREADS] & 0xffu);
                if (P > 0) {
                    const int src = c - P - 1 >= 0 ? (int)(colrec[(c - P - 1) * TINY_THREADS] >> 8) - go : -1;
                    G = G - ge > src ? G - ge : src;
                    if (G < 0) G = 0;
                    int D = G;
                    const int dmax = P < c ? P : c;
                    for (int d = 1; d <= dmax; ++d) { const int v = (int)(colrec[(c - d) * TINY_THREADS] >> 8); D = v > D ? v : D; }
                    mc = D > mc ? D : mc;
                }
                if ((c < e1 || c >= e2) && mc > bv) { bv = mc; bi = c; }
            }
            score2 = bv; ref2 = bv > 0 ? bi : 0;
        }
        int refBegin = -1, readBegin = -1;
        const bool wantBegin = !(a.sc.flag == 0 || (a.sc.flag == 2 && f.score < a.sc.filters));       // ssw.c:834
        if (wantBegin) {
            if (f.score <= 0) { refBegin = -1; readBegin = endRead; }
            else {
                // reversed read prefix [0, endRead] against ref[0, endRef] walked right-to-left (ssw.c:836-849)
                const int pm = endRead + 1, pn = endRef + 1;
                for (int i = 0; i < pm; ++i) { unsigned c = (unsigned char)qb[endRead - i]; qcode[i * TINY_THREADS] = (unsigned char)(c > 4u ? 4u : c); }
                const TinyBest rv = tiny_pass<true>(qb + endRead, -1, pm, rb + endRef, -1, pn, matS, go, ge, HE, colrec, qcode, f.score);
                const int rrow = rv.row > pm - 1 ? pm - 1 : rv.row;
                refBegin = rv.score > 0 ? endRef - rv.col : -1;
                readBegin = endRead - (rv.score > 0 ? rrow : 0);
            }
        }
        rec->score1 = f.score; rec->score2 = score2;
        rec->ref_begin1 = refBegin; rec->ref_end1 = endRef;
        rec->read_begin1 = readBegin; rec->read_end1 = endRead;
        rec->ref_end2 = ref2; rec->cigar_len = 0; rec->cigar_off = 0;
        rec->word = 0;
        rec->status = PS_REV_DONE;
    }
}

cudaError_t launch_tiny(const TinyArgs& a, int blocks, cudaStream_t st)
{
    const int smem = 2 * TINY_LEN * TINY_THREADS * 2 + TINY_LEN * TINY_THREADS;
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        if (dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    tiny_kernel<<<blocks, TINY_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sswb
