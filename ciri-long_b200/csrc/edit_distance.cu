// edit_distance.cu -- batched unit-cost global edit distance.
//
// Replaces CIRI_long/utils.py:153-159 `distance(x, y)` (python-Levenshtein for short strings, edlib in NW
// mode otherwise: the same quantity) as it is called once per pair in collapse.avg_score
// (collapse.py:156-158, 20-nt junction vs the aligned piece of the consensus) and in the O(k^2) loop of
// collapse.cluster_sequence (collapse.py:466-473, homopolymer-compressed reads of up to a few kb).
//
// Bit-parallel Myers / Hyyro recurrences on vertical delta vectors: one 32- or 64-row word of the shorter
// string (the pattern) advances one column of the longer string (the text) in ~17 logic/add instructions,
// the +1/0/-1 horizontal delta of its last row is the carry into the word below.
//   * pattern <= 32 / <= 64 symbols (the junction pairs, millions per run): one THREAD per pair, one word;
//   * longer patterns: G = 4..32 lanes per pair, lane g owns word g of the tile and runs column s-g at
//     step s (the same systolic wavefront as the score kernels): the symbol and the horizontal delta
//     move down one lane per step with a width-G shuffle;
//   * patterns longer than 32 lanes x 32 rows run in row tiles chained through a per-warp array of
//     horizontal deltas (one byte per column).
// Symbols are compared for equality only.  The host maps the bytes that occur in the batch to dense
// 4-bit codes (at most 16 distinct symbols: DNA in both cases plus N), the match masks Peq[symbol] of a
// word live in shared memory, bank == lane.
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

// ---- one thread per pair, one word -----------------------------------------------------------------
template <typename W>
__global__ void __launch_bounds__(EDIT_THREADS) edit_thread_kernel(const EditArgs a)
{
    __shared__ W peq[ED_MAXSYM][EDIT_THREADS];
    __shared__ unsigned char code[256];
    for (int k = threadIdx.x; k < 256; k += EDIT_THREADS) code[k] = a.code[k];
    __syncthreads();
    const int tid = threadIdx.x;
    for (int idx = blockIdx.x * EDIT_THREADS + tid; idx < a.count; idx += gridDim.x * EDIT_THREADS) {
        const int p = a.idx[idx];
        int m = a.x_len[p], n = a.y_len[p];
        const unsigned char* pat = a.seqs + a.x_off[p];
        const unsigned char* txt = a.seqs + a.y_off[p];
        if (m > n) { const int t = m; m = n; n = t; const unsigned char* q = pat; pat = txt; txt = q; }
        if (m == 0) { a.out[p] = n; continue; }
#pragma unroll
        for (int s = 0; s < ED_MAXSYM; ++s) peq[s][tid] = 0;
        for (int i = 0; i < m; ++i) peq[code[pat[i]]][tid] |= (W)1 << i;
        W Pv = ~(W)0, Mv = 0;
        const W top = (W)1 << (m - 1);
        int score = m;
        for (int j = 0; j < n; ++j) {
            const W Eq = peq[code[txt[j]]][tid];
            const W Xv = Eq | Mv;
            const W Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
            W Ph = Mv | ~(Xh | Pv);
            W Mh = Pv & Xh;
            score += (Ph & top) ? 1 : 0;
            score -= (Mh & top) ? 1 : 0;
            Ph = (Ph << 1) | 1;                                  // global distance: row 0 grows by one per column
            Mh <<= 1;
            Pv = Mh | ~(Xv | Ph);
            Mv = Ph & Xv;
        }
        a.out[p] = score;
    }
}

// ---- G lanes per pair, 32 rows per lane and tile ------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(EDIT_THREADS) edit_group_kernel(const EditArgs a)
{
    constexpr int WARPS = EDIT_THREADS / 32;
    __shared__ unsigned peq[WARPS][ED_MAXSYM][32];
    __shared__ unsigned char code[256];
    for (int k = threadIdx.x; k < 256; k += EDIT_THREADS) code[k] = a.code[k];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane % G;
    constexpr int GROUPS = 32 / G;
    const int group = (blockIdx.x * WARPS + warp) * GROUPS + lane / G;
    const int nGroups = gridDim.x * WARPS * GROUPS;
    signed char* carry = a.carry ? a.carry + (size_t)(blockIdx.x * WARPS + warp) * a.carry_stride : nullptr;
    // every group of the warp walks its own pairs; the warp iterates until its slowest group is done
    const int rounds = (a.count + nGroups - 1) / nGroups;
    for (int round = 0; round < rounds; ++round) {
        const int idx = round * nGroups + group;
        const bool have = idx < a.count;
        const int p = have ? a.idx[idx] : 0;
        int m = have ? a.x_len[p] : 0, n = have ? a.y_len[p] : 0;
        const unsigned char* pat = a.seqs + (have ? a.x_off[p] : 0);
        const unsigned char* txt = a.seqs + (have ? a.y_off[p] : 0);
        if (m > n) { const int t = m; m = n; n = t; const unsigned char* q = pat; pat = txt; txt = q; }
        const int tiles = (m + 32 * G - 1) / (32 * G);
        int maxTiles = tiles, maxSteps = n + G;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            maxTiles = max(maxTiles, __shfl_xor_sync(FULL, maxTiles, o));
            maxSteps = max(maxSteps, __shfl_xor_sync(FULL, maxSteps, o));
        }
        int score = m;
        for (int t = 0; t < maxTiles; ++t) {
            const int base = (t * G + g) * 32;                    // first pattern row of my word
            const bool mine = t < tiles && base < m;
            const bool lastTile = t == tiles - 1;
            __syncwarp();
#pragma unroll
            for (int s = 0; s < ED_MAXSYM; ++s) peq[warp][s][lane] = 0;
            int rows = 0;
            if (mine) {
                rows = m - base < 32 ? m - base : 32;
                for (int i = 0; i < rows; ++i) peq[warp][code[pat[base + i]]][lane] |= 1u << i;
            }
            // the word that holds the tile's last row hands its delta to the next tile (or to the score)
            const bool tail = mine && (base + 32 >= m || g == G - 1);
            const unsigned topBit = rows ? rows - 1 : 0;
            const unsigned* myPeq = &peq[warp][0][lane];
            unsigned Pv = 0xffffffffu, Mv = 0;
            // horizontal deltas travel as two bits: bit 0 = +1, bit 1 = -1
            unsigned hout = 0, c = 0, symv = 0, carv = 1;
            for (int s = 0; s < maxSteps; ++s) {
                if ((s & (G - 1)) == 0) {
                    // every G steps the group fetches the next G text symbols (and, below the first tile,
                    // the deltas the tile above left for these columns), one per lane
                    const int col = s + g;
                    symv = col < n ? code[txt[col]] : 0;
                    carv = (t > 0 && col < n) ? (unsigned)carry[col] : 1u;
                }
                unsigned cin = __shfl_up_sync(FULL, c, 1, G);
                unsigned hin = __shfl_up_sync(FULL, hout, 1, G);
                const unsigned c0 = __shfl_sync(FULL, symv, s & (G - 1), G);
                const unsigned h0 = __shfl_sync(FULL, carv, s & (G - 1), G);
                if (g == 0) { cin = c0; hin = h0; }
                c = cin;
                const int j = s - g;
                if ((unsigned)j < (unsigned)n) {
                    const unsigned hp = hin & 1u, hm = hin >> 1;
                    unsigned Eq = myPeq[c * 32];
                    const unsigned Xv = Eq | Mv;
                    Eq |= hm;
                    const unsigned Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                    unsigned Ph = Mv | ~(Xh | Pv);
                    unsigned Mh = Pv & Xh;
                    hout = ((Ph >> topBit) & 1u) | (((Mh >> topBit) & 1u) << 1);
                    Ph = (Ph << 1) | hp;
                    Mh = (Mh << 1) | hm;
                    Pv = Mh | ~(Xv | Ph);
                    Mv = Ph & Xv;
                    if (tail) {
                        if (lastTile) score += (int)(hout & 1u) - (int)(hout >> 1);
                        else carry[j] = (signed char)hout;
                    }
                }
            }
        }
        // the lane that owns the pattern's last row holds the distance
        if (have && m > 0) {
            const int owner = ((m - 1) / 32) % G;
            if (g == owner) a.out[p] = score;
        } else if (have && g == 0) a.out[p] = n;
    }
}

cudaError_t launch_edit(int kind, const EditArgs& a, int blocks, cudaStream_t st)
{
    switch (kind) {
        case 0: edit_thread_kernel<unsigned><<<blocks, EDIT_THREADS, 0, st>>>(a); break;
        case 1: edit_thread_kernel<unsigned long long><<<blocks, EDIT_THREADS, 0, st>>>(a); break;
        case 2: edit_group_kernel<4><<<blocks, EDIT_THREADS, 0, st>>>(a); break;
        case 3: edit_group_kernel<8><<<blocks, EDIT_THREADS, 0, st>>>(a); break;
        case 4: edit_group_kernel<16><<<blocks, EDIT_THREADS, 0, st>>>(a); break;
        default: edit_group_kernel<32><<<blocks, EDIT_THREADS, 0, st>>>(a); break;
    }
    return cudaGetLastError();
}

}  // namespace sswb
