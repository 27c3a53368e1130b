// ssw_lists.cu -- device-side work lists.  Every stage of the hot path (forward pass, deciding pass,
// reverse pass, CIGAR pass) consumes a list of pair indices grouped by the kernel instance that must
// process them (strip height K, recurrence flavour, long-reference class), so that no stage needs a
// host round trip: counts and cursors live in device memory and the consumers read them there.
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

// rows per strip: GOTOH cuts the query into tiles of 64 equal strips; TRUNC gives each of the reference's
// 8 segments (ceil(m/8) rows) its own G strips (ssw_score_impl.cuh)
__device__ __forceinline__ int strip_height(int m, int kind) { return strip_height_for(m, kind); }

// which list a pair belongs to in `stage`; -1 = not part of this stage
__device__ int classify(int stage, const BatchView& b, const Scoring& sc, int long_thr, int p, int maxScore)
{
    PairRec* rec = b.rec + p;
    if (stage == 0) {
        const int m = b.q_len[p], n = b.r_len[p];
        if (m <= 0 || n <= 0) return -1;
        // the word flavour with gap_open == gap_extend has its own recurrence (TRUNC); which one goes first is a guess
        if (is_tiny_pair(m, n, maxScore, sc.bias)) return LIST_TINY;
        const int kind = first_pass_kind(m, sc.go, sc.ge, maxScore, sc.bias);
        return list_id(n > long_thr ? 1 : 0, kind, strip_height(m, kind));
    }
    // stage 1: reverse pass (ssw.c:834)
    if (rec->status & (PS_PUNT | PS_UNSUPPORTED | PS_REV_DONE)) return -1;
    if (sc.flag == 0 || (sc.flag == 2 && rec->score1 < sc.filters)) return -1;
    const int m = rec->read_end1 + 1, n = rec->ref_end1 + 1;
    if (n <= 0) return -2;          // score 0: nothing to walk (handled inline by the caller)
    const int kind = (rec->word && sc.go == sc.ge) ? 1 : 0;
    if (rec->status & PS_WIDE32) return LIST_WIDE32 + kind;
    // the scratch class follows the full reference length (as in the forward pass), not the trimmed one
    return list_id(b.r_len[p] > long_thr ? 1 : 0, kind, strip_height(m, kind));
}

__device__ __forceinline__ int max_score(const Scoring& sc)
{
    int mx = 0;
    for (int k = 0; k < 25; ++k) mx = sc.mat[k] > mx ? sc.mat[k] : mx;
    return mx;
}

__global__ void list_count_kernel(int stage, BatchView b, Scoring sc, int long_thr, ListSet ls)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int maxScore = max_score(sc);
    int id = -1;
    if (p < b.n_pairs) {
        id = classify(stage, b, sc, long_thr, p, maxScore);
        PairRec* rec = b.rec + p;
        if (stage == 0) {
            // default record (also the answer for empty inputs)
            rec->score1 = 0; rec->score2 = 0; rec->ref_begin1 = -1; rec->ref_end1 = -1;
            rec->read_begin1 = -1; rec->read_end1 = -1; rec->ref_end2 = -1;
            rec->cigar_len = 0; rec->cigar_off = 0; rec->word = 0;
            rec->status = id < 0 ? PS_UNSUPPORTED : 0;
        } else if (id == -2) {
            rec->ref_begin1 = -1;                       // byte flavour, end_ref stays -1 (ssw.c:145)
            rec->read_begin1 = rec->read_end1;
        }
    }
    const unsigned act = __ballot_sync(0xffffffffu, id >= 0);
    if (id >= 0) {
        const unsigned peers = __match_any_sync(act, id);
        if ((int)(__ffs(peers) - 1) == lane_id()) atomicAdd(&ls.count[id], __popc(peers));
    }
}

__global__ void list_scan_kernel(ListSet ls)
{
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < N_LISTS; ++k) { ls.base[k] = run; run += ls.count[k]; ls.fill[k] = 0; ls.cursor[k] = 0; }
    }
}

__global__ void list_scatter_kernel(int stage, BatchView b, Scoring sc, int long_thr, ListSet ls)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int maxScore = max_score(sc);
    int id = -1;
    if (p < b.n_pairs) id = classify(stage, b, sc, long_thr, p, maxScore);
    const unsigned act = __ballot_sync(0xffffffffu, id >= 0);
    if (id >= 0) {
        const unsigned peers = __match_any_sync(act, id);
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        if (lane_id() == leader) pos = atomicAdd(&ls.fill[id], __popc(peers));
        pos = __shfl_sync(peers, pos, leader);
        ls.idx[ls.base[id] + pos + __popc(peers & ((1u << lane_id()) - 1u))] = p;
    }
}

cudaError_t build_lists(int stage, const BatchView& b, const Scoring& sc, int long_thr, const ListSet& ls,
                        cudaStream_t st, int* launches)
{
    cudaError_t e = cudaMemsetAsync(ls.count, 0, N_LISTS * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    const int threads = 256, blocks = (b.n_pairs + threads - 1) / threads;
    list_count_kernel<<<blocks, threads, 0, st>>>(stage, b, sc, long_thr, ls);
    list_scan_kernel<<<1, 32, 0, st>>>(ls);
    list_scatter_kernel<<<blocks, threads, 0, st>>>(stage, b, sc, long_thr, ls);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// ---- CIGAR stage: one list with every pair that passes the reference's gate (ssw.c:850)
__global__ void band_list_kernel(BatchView b, Scoring sc, ListSet ls)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    if (p < b.n_pairs) {
        const PairRec* rec = b.rec + p;
        const int flag = sc.flag;
        want = !(rec->status & (PS_PUNT | PS_UNSUPPORTED)) &&
               !(flag == 0 || (flag == 2 && rec->score1 < sc.filters)) &&
               !((7 & flag) == 0 || ((2 & flag) != 0 && rec->score1 < sc.filters) ||
                 ((4 & flag) != 0 && (rec->ref_end1 - rec->ref_begin1 > sc.filterd ||
                                      rec->read_end1 - rec->read_begin1 > sc.filterd)));
    }
    const unsigned peers = __ballot_sync(0xffffffffu, want);
    if (want) {
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        if (lane_id() == leader) pos = atomicAdd(&ls.count[0], __popc(peers));
        pos = __shfl_sync(peers, pos, leader);
        ls.idx[pos + __popc(peers & ((1u << lane_id()) - 1u))] = p;
    }
}

cudaError_t build_band_list(const BatchView& b, const Scoring& sc, const ListSet& ls, cudaStream_t st, int* launches)
{
    cudaError_t e = cudaMemsetAsync(ls.count, 0, N_LISTS * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(ls.cursor, 0, N_LISTS * sizeof(int32_t), st);
    if (e != cudaSuccess) return e;
    const int threads = 256, blocks = (b.n_pairs + threads - 1) / threads;
    band_list_kernel<<<blocks, threads, 0, st>>>(b, sc, ls);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// ---- long references: one task per column chunk (ssw_kernels.h: ChunkPlan)
__global__ void expand_tasks_kernel(WorkList wl, int rev, BatchView b, Scoring sc, ChunkPlan ck, int32_t* task_count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *wl.count) return;
    const int pair = wl.idx[(wl.base ? *wl.base : 0) + i];
    // forward: the whole pair; reverse: the prefix rectangle that ends at the forward pass's best cell
    const int m = rev ? b.rec[pair].read_end1 + 1 : b.q_len[pair];
    const int n = rev ? b.rec[pair].ref_end1 + 1 : b.r_len[pair];
    const int nt = chunk_tasks(m, n, ck.chunk_cols, ck.max_match, sc.ge);
    const int at = atomicAdd(task_count, nt);
    for (int t = 0; t < nt; ++t) {
        const int c0 = nt == 1 ? 0 : t * ck.chunk_cols;
        ck.task_pair[at + t] = pair;
        ck.task_c0[at + t] = c0;
        ck.task_c1[at + t] = (nt == 1 || c0 + ck.chunk_cols > n) ? n : c0 + ck.chunk_cols;
    }
    ck.pair_left[pair] = nt;
    ck.pair_key[pair] = rev ? ((unsigned long long)at << 32) | (unsigned)nt : 0ull;
}

cudaError_t expand_tasks(const WorkList& wl, int max_pairs, bool rev, const BatchView& b, const Scoring& sc, const ChunkPlan& ck,
                         int32_t* task_count, cudaStream_t st, int* launches)
{
    if (max_pairs <= 0) return cudaSuccess;
    expand_tasks_kernel<<<(max_pairs + 255) / 256, 256, 0, st>>>(wl, rev ? 1 : 0, b, sc, ck, task_count);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// ---- ASCII -> {0..4} on the device: the host uploads the raw letters and this
// pass replaces the per-base encode of ssw_wrap.py:234-252 (A C G T N in either case, anything else N)
__global__ void encode_ascii_kernel(int8_t* seqs, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x * 16;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n; i += stride) {
        // 16 bytes per thread and step when the tail allows it
        if (i + 16 <= n && ((reinterpret_cast<uintptr_t>(seqs + i) & 15) == 0)) {
            uint4 v = *reinterpret_cast<uint4*>(seqs + i);
            unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned o = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const unsigned c = ((w[k] >> (8 * b)) & 0xffu) | 0x20u;         // lower case
                    const unsigned code = c == 'a' ? 0u : c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : 4u;
                    o |= code << (8 * b);
                }
                w[k] = o;
            }
            *reinterpret_cast<uint4*>(seqs + i) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (long long j = i; j < n && j < i + 16; ++j) {
                const unsigned c = (unsigned)(unsigned char)seqs[j] | 0x20u;
                seqs[j] = (int8_t)(c == 'a' ? 0 : c == 'c' ? 1 : c == 'g' ? 2 : c == 't' ? 3 : 4);
            }
        }
    }
}

cudaError_t encode_ascii(int8_t* seqs, long long n, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    long long blocks = (n / 16 + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    if (blocks < 1) blocks = 1;
    encode_ascii_kernel<<<(int)blocks, 256, 0, st>>>(seqs, n);
    return cudaGetLastError();
}

// ---- 4-bit packed bases (two per byte, low nibble first) -> one code per byte; nibbles above 4 count as N.
// 16 packed bytes in, 32 codes out per thread and step: 128-bit loads and stores.
__global__ void unpack4_kernel(const unsigned char* __restrict__ packed, int8_t* __restrict__ codes, long long n_bytes)
{
    const long long stride = (long long)gridDim.x * blockDim.x * 16;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < n_bytes; i += stride) {
        if (i + 16 <= n_bytes) {
            const uint4 v = *reinterpret_cast<const uint4*>(packed + i);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
            unsigned o[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // bytes b3 b2 b1 b0 -> nibbles: low word = (b1.hi b1.lo b0.hi b0.lo), high word = (b3.hi b3.lo b2.hi b2.lo)
                unsigned lo = (w[k] & 0xfu) | ((w[k] & 0xf0u) << 4) | ((w[k] & 0xf00u) << 8) | ((w[k] & 0xf000u) << 12);
                unsigned hi = ((w[k] >> 16) & 0xfu) | (((w[k] >> 16) & 0xf0u) << 4) | (((w[k] >> 16) & 0xf00u) << 8) | (((w[k] >> 16) & 0xf000u) << 12);
                o[2 * k] = __vminu4(lo, 0x04040404u);
                o[2 * k + 1] = __vminu4(hi, 0x04040404u);
            }
            reinterpret_cast<uint4*>(codes + 2 * i)[0] = make_uint4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<uint4*>(codes + 2 * i)[1] = make_uint4(o[4], o[5], o[6], o[7]);
        } else {
            for (long long j = i; j < n_bytes; ++j) {
                const unsigned b = packed[j];
                codes[2 * j] = (int8_t)min(b & 0xfu, 4u);
                codes[2 * j + 1] = (int8_t)min(b >> 4, 4u);
            }
        }
    }
}

cudaError_t unpack4(const unsigned char* packed, int8_t* codes, long long n_bytes, cudaStream_t st)
{
    if (n_bytes <= 0) return cudaSuccess;
    long long blocks = (n_bytes / 16 + 255) / 256;
    if (blocks > 8192) blocks = 8192;
    if (blocks < 1) blocks = 1;
    unpack4_kernel<<<(int)blocks, 256, 0, st>>>(packed, codes, n_bytes);
    return cudaGetLastError();
}

// A list that is still being appended to, consumed in two launches: the first launch works on a snapshot of the
// count (its warps overshoot the fetch cursor when they leave), then the cursor is put back on the snapshot so
// that the second launch continues exactly where the first one ended.
__global__ void copy_int_kernel(int32_t* dst, const int32_t* src) { *dst = *src; }
cudaError_t snapshot_count(int32_t* snap, const int32_t* count, cudaStream_t st) { copy_int_kernel<<<1, 1, 0, st>>>(snap, count); return cudaGetLastError(); }
cudaError_t rewind_cursor(int32_t* cursor, const int32_t* snap, cudaStream_t st) { copy_int_kernel<<<1, 1, 0, st>>>(cursor, snap); return cudaGetLastError(); }

__global__ void clear_status_kernel(BatchView b, int bits)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < b.n_pairs) { b.rec[p].status &= ~bits; b.rec[p].cigar_len = 0; b.rec[p].cigar_off = 0; }
}

cudaError_t clear_status_bits(const BatchView& b, int bits, cudaStream_t st)
{
    clear_status_kernel<<<(b.n_pairs + 255) / 256, 256, 0, st>>>(b, bits);
    return cudaGetLastError();
}

}  // namespace sswb
