// ssw_tband.cu -- throughput CIGAR pass: banded DP + traceback with one pair per LANE (ssw_tband_core.h).
//
// Replaces banded_sw (reference ssw.c:548-735) as called from ssw_align (ssw.c:852-856) for large batches;
// ssw_band.cu (one pair per warp) stays the path for small batches and for the pairs this kernel hands over
// (bands wider than 126, scores near the 16-bit range, score-0 pairs, staging overflow).
//
// Pipeline per band-doubling pass p (ssw.c:571-632: the do-loop, one iteration per pass):
//   1. tband_key_kernel     bin every pair of the pass's list by (blocks of 4 steps its band needs, row pairs),
//   2. tband_scan_kernel    prefix sums over the bins, largest work first; segment borders per instance,
//   3. tband_scatter_kernel sorted list,
//   4. tband_kernel         x instances (shared memory sized for bands of <= 8/16/32/64 blocks): persistent
//                           warps take 32 consecutive pairs of the sorted list (similar band and length, so
//                           lock step costs little), run the fill in lock step, then every lane whose pass is
//                           final walks its own traceback; the others go to the next pass's list.
// Direction words go to a per-warp scratch in HBM, laid out [row pair][block][lane]: every store of the fill is
// one full 128-byte line per warp; the traceback reads 4 bytes per step and lane.
#include <cuda_runtime.h>
#include <functional>
#include "ssw_common.cuh"
#include "ssw_kernels.h"
#include "ssw_tband_core.h"

namespace sswb {
using namespace sswt;

constexpr int TBAND_WARPS = 4;                       // warps per CTA
constexpr int TB_NB_MAX = TB_MAX_STEPS / TB_BLOCK;   // 32 blocks
constexpr int TB_INST = 5;                           // shared-memory instances: bands of up to 8/12/16/24/32 blocks
constexpr int TB_PASSES = 5;                         // band-doubling passes run here; what is left goes to ssw_band.cu
constexpr int TB_RB_SHIFT = 3, TB_RB_MAX = 255;      // row-pair bins of 8 row pairs
constexpr int TB_BINS = TB_NB_MAX * (TB_RB_MAX + 1);

__device__ __forceinline__ int tb_bin(int nb, int rowPairs)
{
    int rb = rowPairs >> TB_RB_SHIFT; if (rb > TB_RB_MAX) rb = TB_RB_MAX;
    return (TB_NB_MAX - nb) * (TB_RB_MAX + 1) + (TB_RB_MAX - rb);            // widest band, then most rows, first
}
__host__ __device__ inline int tb_instance_of(int nb) { return nb <= 8 ? 0 : nb <= 12 ? 1 : nb <= 16 ? 2 : nb <= 24 ? 3 : 4; }
__host__ __device__ inline int tb_instance_nb(int inst) { return inst == 0 ? 8 : inst == 1 ? 12 : inst == 2 ? 16 : inst == 3 ? 24 : 32; }
// A launch costs at least one lock-step round of 32 pairs per warp (milliseconds for wide bands), so segments
// with fewer pairs than this are cheaper in the warp-per-pair instance.
__host__ __device__ inline int tb_instance_min_pairs(int inst) { return inst == 0 ? 3072 : inst == 1 ? 4096 : inst == 2 ? 6144 : inst == 3 ? 10240 : 16384; }

// Geometry of one pair in pass `pass`: band width, rows.  Returns false if the pair leaves for ssw_band.cu.
__device__ __forceinline__ bool tb_geometry(const TbandArgs& a, const PairRec* rec, int pass, int& bw, int& readLen, int& refLen)
{
    if (rec->ref_begin1 < 0) return false;                                   // score 0: "1M" (ssw_band.cu)
    refLen = rec->ref_end1 - rec->ref_begin1 + 1;
    readLen = rec->read_end1 - rec->read_begin1 + 1;
    if (pass == 0) { bw = refLen - readLen; if (bw < 0) bw = -bw; bw += 1; }
    else bw = -rec->cigar_len;
    if (tb_steps(bw) > TB_MAX_STEPS) return false;
    if ((readLen + 1) / 2 > a.row_pairs_cap) return false;
    if (rec->score1 + a.sc.go + a.sc.ge >= TB_SCORE_LIMIT) return false;
    return true;
}

__device__ __forceinline__ void tb_hand_over(const TbandArgs& a, PairRec* rec, int pair, int bw, int maxv)
{
    rec->cigar_len = -bw; rec->cigar_off = maxv;                             // ssw_band.cu: band_pair<CLASS > 0> starts from these
    if (2 * bw + 1 <= 256) a.fallback1_idx[atomicAdd(a.fallback1_count, 1)] = pair;   // class 1: up to 256 diagonals
    else a.fallback_idx[atomicAdd(a.fallback_count, 1)] = pair;
}

__global__ void tband_key_kernel(TbandArgs a, int pass, const int32_t* in_idx, const int32_t* in_count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    int bin = -1;
    if (k < *in_count) {
        const int pair = in_idx[k];
        PairRec* rec = a.b.rec + pair;
        int bw, readLen, refLen;
        if (!tb_geometry(a, rec, pass, bw, readLen, refLen)) {
            if (pass == 0) { int w = 1; if (rec->ref_begin1 >= 0) { w = rec->ref_end1 - rec->ref_begin1 - (rec->read_end1 - rec->read_begin1); if (w < 0) w = -w; w += 1; } tb_hand_over(a, rec, pair, w, 0); }
            else tb_hand_over(a, rec, pair, -rec->cigar_len, (int)rec->cigar_off);
        } else {
            if (pass == 0) { rec->cigar_len = -bw; rec->cigar_off = 0; }
            bin = tb_bin(tb_blocks(bw), (readLen + 1) / 2);
        }
        a.keys[k] = bin;
    }
    // one atomic per distinct bin and warp: batches of alike pairs (millions of junction pairs in a handful of bins)
    // would otherwise serialise on a few counters
    const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
    if (bin >= 0) {
        const unsigned peers = __match_any_sync(act, bin);
        if ((int)(__ffs(peers) - 1) == lane_id()) atomicAdd(a.bin_count + bin, __popc(peers));
    }
}

// one CTA: exclusive prefix sums over the bins (bin order = processing order), per-instance segments
__global__ void tband_scan_kernel(TbandArgs a)
{
    __shared__ int part[1024];
    const int per = (TB_BINS + 1023) / 1024;
    const int t = threadIdx.x;
    int sum = 0;
    for (int k = t * per; k < (t + 1) * per && k < TB_BINS; ++k) sum += a.bin_count[k];
    part[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - sum;
    for (int k = t * per; k < (t + 1) * per && k < TB_BINS; ++k) {
        const int c = a.bin_count[k];
        a.bin_base[k] = run; a.bin_count[k] = 0;                             // count becomes the scatter cursor / is clean for the next pass
        run += c;
    }
    __syncthreads();
    if (t < TB_INST) {
        // instance t covers bands of (nbLo, nbHi] blocks = bins [(NB_MAX - nbHi) * R, (NB_MAX - nbLo) * R)
        const int nbHi = tb_instance_nb(t), nbLo = t == 0 ? 0 : tb_instance_nb(t - 1);
        const int b0 = (TB_NB_MAX - nbHi) * (TB_RB_MAX + 1), b1 = (TB_NB_MAX - nbLo) * (TB_RB_MAX + 1);
        const int lo = a.bin_base[b0];
        const int hi = b1 < TB_BINS ? a.bin_base[b1] : part[1023];
        const bool small = hi - lo < a.min_pairs_scale * tb_instance_min_pairs(t) / 16;
        a.seg[t] = lo; a.seg[8 + t] = small ? 0 : hi - lo; a.seg[16 + t] = 0; a.seg[24 + t] = small ? 1 : 0;   // base, count, cursor, handed over
    }
}

__global__ void tband_scatter_kernel(TbandArgs a, const int32_t* in_idx, const int32_t* in_count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    int bin = -1, pair = 0;
    if (k < *in_count) {
        bin = a.keys[k];
        pair = in_idx[k];
        if (bin >= 0 && a.seg[24 + tb_instance_of(TB_NB_MAX - bin / (TB_RB_MAX + 1))]) {        // too few pairs for a launch of that instance
            PairRec* rec = a.b.rec + pair;
            tb_hand_over(a, rec, pair, -rec->cigar_len, (int)rec->cigar_off);
            bin = -1;
        }
    }
    const unsigned act = __ballot_sync(0xffffffffu, bin >= 0);
    if (bin >= 0) {
        const unsigned peers = __match_any_sync(act, bin);
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        if (lane_id() == leader) pos = atomicAdd(a.bin_count + bin, __popc(peers));
        pos = __shfl_sync(peers, pos, leader);
        a.sorted[a.bin_base[bin] + pos + __popc(peers & ((1u << lane_id()) - 1u))] = pair;
    }
}

__global__ void tband_reset_kernel(TbandArgs a, int32_t* count_to_clear)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < TB_BINS) a.bin_count[k] = 0;
    if (k == 0 && count_to_clear) *count_to_clear = 0;
}

// shared memory of one warp: S (H|E of the stored row), reference ring (bytes), score tables
__host__ __device__ inline int tband_warp_smem(int nbcap) { return (4 * nbcap + 4) * TB_LANES * 4 + (4 * nbcap + 4) * TB_LANES + 10 * TB_LANES * 4; }

__global__ void __launch_bounds__(TBAND_WARPS * 32) tband_kernel(const TbandArgs a, const int inst, const int nbcap, const int pass)
{
    extern __shared__ unsigned tb_smem[];
    __shared__ int matS[25];
    const int count = a.seg[8 + inst];
    if (count <= 0) return;
    if (threadIdx.x < 25) matS[threadIdx.x] = a.sc.mat[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int segBase = a.seg[inst];
    unsigned char* wsm = reinterpret_cast<unsigned char*>(tb_smem) + (size_t)warp * tband_warp_smem(nbcap);
    unsigned* S = reinterpret_cast<unsigned*>(wsm) + lane;
    unsigned char* ring = wsm + (4 * nbcap + 4) * TB_LANES * 4 + lane;
    int* tab = reinterpret_cast<int*>(wsm + (4 * nbcap + 4) * TB_LANES * 5) + lane;
    const unsigned tabS = (unsigned)__cvta_generic_to_shared(tab);
    unsigned char* wscr = a.scratch + (size_t)(blockIdx.x * TBAND_WARPS + warp) * a.scratch_stride;
    unsigned* dirs = reinterpret_cast<unsigned*>(wscr) + lane;
    unsigned* stage = reinterpret_cast<unsigned*>(wscr + a.dir_bytes) + lane;

    const int go = a.sc.go, ge = a.sc.ge, B = go + ge;
    const unsigned B2 = (unsigned)B * 0x10001u, GO2 = (unsigned)go * 0x10001u, GE2 = (unsigned)ge * 0x10001u;
    const unsigned one = (unsigned)a.one;                 // 1, but not a compile-time constant (ssw_tband_core.h: tb_max2_*_wins)

    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(a.seg + 16 + inst, 32);
        base = __shfl_sync(FULL, base, 0);
        if (base >= count) break;
        const bool have = base + lane < count;
        const int pair = a.sorted[segBase + (have ? base + lane : count - 1)];   // idle lanes shadow the last pair, outputs off
        PairRec* rec = a.b.rec + pair;
        TbJob J;
        {
            int bw, readLen, refLen;
            tb_geometry(a, rec, 1, bw, readLen, refLen);                     // (band width was stored by the key kernel)
            J.ref = a.b.seqs + a.b.r_off[pair] + rec->ref_begin1;
            J.read = a.b.seqs + a.b.q_off[pair] + rec->read_begin1;
            J.refLen = refLen; J.readLen = readLen; J.bw = bw;
            J.score = rec->score1; J.maxIn = (int)rec->cigar_off;
        }
        const int myRowPairs = (J.readLen + 1) / 2;
        const int NB = __reduce_max_sync(FULL, tb_blocks(J.bw));
        const int rowPairsMax = __reduce_max_sync(FULL, myRowPairs);
        const int RING = 4 * NB;
        auto refcode = [&](int col) -> unsigned { return tb_ref_code(J, col); };
        for (int k = 0; k < 4 * NB + 4; ++k) S[k * TB_LANES] = B2;
        for (int p = 0; p < RING; ++p) ring[p * TB_LANES] = (unsigned char)refcode(p - J.bw);
        for (int p = 0; p < 4; ++p) ring[(RING + p) * TB_LANES] = ring[p * TB_LANES];
        unsigned maxv2 = (unsigned)(J.maxIn + B) * 0x10001u, finalMax = maxv2;
        __syncwarp();

        TbRow R;
        int pos0row = 0;
        unsigned* drow = dirs;
        unsigned nr0 = tb_read_code(J, 0), nr1 = tb_read_code(J, 1);
        for (int rho = 0; rho < rowPairsMax; ++rho) {
            // loads of the next row pair's read bases and of the two reference bases that enter the window after
            // this row pair are issued here and consumed at its end
            const unsigned r0 = nr0, r1 = nr1;
            nr0 = tb_read_code(J, 2 * rho + 2); nr1 = tb_read_code(J, 2 * rho + 3);
            const unsigned cNew0 = refcode(2 * rho + RING - J.bw), cNew1 = refcode(2 * rho + RING + 1 - J.bw);
            tb_row_begin(R, J, rho, S, tab, tabS, matS, B2, r0, r1);
            // which blocks need which body (ssw_tband_core.h: tb_row_plan), agreed over the lanes that still have rows
            TbRowPlan pl = tb_row_plan(R, NB);
            if (rho >= myRowPairs) { pl.lo = 0; pl.hi = 4 * NB; pl.simple = true; }
            const int head = __reduce_max_sync(FULL, pl.lo), tail = __reduce_min_sync(FULL, pl.hi + 1);
            const bool simple = __all_sync(FULL, pl.simple);
            int hb = (head + 3) >> 2, tb = tail >> 2;
            if (tb < hb) { hb = NB; tb = NB; }
            int pos = pos0row;
            int b = 0;
            if (simple && hb == 1) {
                drow[0] = tb_block<TB_HEAD>(R, 0, S, ring + pos * TB_LANES, tabS, B2, GO2, GE2, one, maxv2);
                pos += 4; if (pos >= RING) pos -= RING;
                b = 1;
            }
            for (; b < hb; ++b) {
                drow[b * TB_LANES] = tb_block<TB_ANY>(R, 4 * b, S, ring + pos * TB_LANES, tabS, B2, GO2, GE2, one, maxv2);
                pos += 4; if (pos >= RING) pos -= RING;
            }
            for (; b < tb; ++b) {
                drow[b * TB_LANES] = tb_block<TB_PLAIN>(R, 4 * b, S, ring + pos * TB_LANES, tabS, B2, GO2, GE2, one, maxv2);
                pos += 4; if (pos >= RING) pos -= RING;
            }
            for (; b < NB; ++b) {
                drow[b * TB_LANES] = tb_block<TB_TAIL>(R, 4 * b, S, ring + pos * TB_LANES, tabS, B2, GO2, GE2, one, maxv2);
                pos += 4; if (pos >= RING) pos -= RING;
            }
            drow += NB * TB_LANES;
            // the window moves two columns to the right
            ring[pos0row * TB_LANES] = (unsigned char)cNew0;                  // pos0row is even and < RING
            ring[(pos0row + 1) * TB_LANES] = (unsigned char)cNew1;
            if (pos0row < 4) { ring[(RING + pos0row) * TB_LANES] = (unsigned char)cNew0; ring[(RING + pos0row + 1) * TB_LANES] = (unsigned char)cNew1; }
            pos0row += 2; if (pos0row >= RING) pos0row -= RING;
            if (rho + 1 == myRowPairs) finalMax = maxv2;
        }

        // ---- end of the pass (ssw.c:631-632)
        const int mx = (lo16(finalMax) > hi16(finalMax) ? lo16(finalMax) : hi16(finalMax)) - B;
        const bool again = mx < J.score && J.bw < J.readLen;
        int nOps = 0, status = 0;
        bool fallback = false;
        if (have && !again) {
            nOps = tb_traceback(J, dirs, NB, stage, a.stage_cap);
            if (nOps == -1) { status = PS_TRACEBACK_ERR; nOps = 0; }
            else if (nOps == -2) { fallback = true; nOps = 0; }
        }
        // CIGAR space for the whole warp with one atomic
        int incl = nOps;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
        const int total = __shfl_sync(FULL, incl, 31);
        long long wbase = 0;
        if (lane == 31 && total > 0) wbase = (long long)atomicAdd(a.cigar_used, (unsigned long long)total);
        wbase = __shfl_sync(FULL, wbase, 31);
        const long long off = wbase + incl - nOps;
        // pairs whose band doubles: next pass's list, one atomic per warp
        const int bw2 = 2 * J.bw;
        const bool toNext = have && again && !(tb_steps(bw2) > TB_MAX_STEPS || pass + 1 >= TB_PASSES);
        const unsigned nextMask = __ballot_sync(FULL, toNext);
        if (nextMask) {
            int nbase = 0;
            const int leader = __ffs(nextMask) - 1;
            if (lane == leader) nbase = atomicAdd(a.next_count, __popc(nextMask));
            nbase = __shfl_sync(FULL, nbase, leader);
            if (toNext) { rec->cigar_len = -bw2; rec->cigar_off = mx; a.next_idx[nbase + __popc(nextMask & ((1u << lane) - 1u))] = pair; }
        }
        if (have) {
            if (again) {
                if (!toNext) tb_hand_over(a, rec, pair, bw2, mx);
            } else if (fallback) tb_hand_over(a, rec, pair, J.bw, (int)rec->cigar_off);
            else {
                if (nOps > 0) {
                    if (off + nOps <= a.cigar_cap) { for (int k = 0; k < nOps; ++k) a.cigar_buf[off + k] = stage[(nOps - 1 - k) * TB_LANES]; }
                    else { status |= PS_CIGAR_CAP; }
                }
                rec->cigar_off = off;
                rec->cigar_len = (status & (PS_CIGAR_CAP | PS_TRACEBACK_ERR)) ? 0 : nOps;
                rec->status |= status;
            }
        }
        __syncwarp();
    }
}

// ---- host side ------------------------------------------------------------------------------------------
cudaError_t tband_plan(int device, int sms, int max_q, long long budget, TbandPlan* plan)
{
    int rp = (max_q + 1) / 2; if (rp < 1) rp = 1;
    plan->row_pairs_cap = rp;
    plan->stage_cap = 4 * rp + 64;
    int smemMax = 0;
    cudaError_t e = cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return e;
    long long total = 0;
    budget /= TB_INST;                                    // the instances of a pass run side by side, each in its own region
    for (int inst = 0; inst < TB_INST; ++inst) {
        const int nbcap = tb_instance_nb(inst);
        const int smem = TBAND_WARPS * tband_warp_smem(nbcap);
        int perSm = smemMax / (smem + 1024 + 128); if (perSm > 8) perSm = 8; if (perSm < 1) perSm = 1;   // (+1 KB per CTA reserved by the driver)
        const long long dirBytes = (((long long)rp * nbcap * TB_LANES * 4) + 255) & ~255LL;
        const long long stride = dirBytes + (((long long)plan->stage_cap * TB_LANES * 4 + 255) & ~255LL);
        long long blocks = (long long)sms * perSm;
        const long long fit = budget / (stride * TBAND_WARPS);
        if (blocks > fit) blocks = fit;
        if (blocks < 1) blocks = 1;
        plan->blocks[inst] = (int)blocks; plan->smem[inst] = smem; plan->dir_bytes[inst] = dirBytes; plan->stride[inst] = stride;
        plan->scratch_off[inst] = total;
        total += blocks * TBAND_WARPS * stride;
    }
    plan->scratch_bytes = total;
    return cudaSuccess;
}

cudaError_t tband_configure()
{
    int dev = 0, smemMax = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&smemMax, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tband_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax - 1024);   // (static: the 5x5 matrix)
}

// All passes of the CIGAR stage for the pairs of in_idx[0 .. *in_count).  Pairs this kernel cannot take end in the
// two hand-over lists (consumed by launch_band(1 / 2, ...)).  The instances of a pass are independent launches:
// each runs on its own stream between two events, so a launch that is down to its last (long) lock-step rounds
// shares the machine with the others instead of holding it.  `after_first_sort` is called once, when the first
// pass's hand-overs are known (the caller starts the warp-per-pair instance on them, side by side).
cudaError_t launch_tband(TbandArgs a, const TbandPlan& plan, const int32_t* in_idx, const int32_t* in_count, int n_max,
                         int32_t* list_a, int32_t* list_b, int32_t* cnt_a, int32_t* cnt_b, cudaStream_t st,
                         cudaStream_t* side, cudaEvent_t* ev, const std::function<cudaError_t(cudaEvent_t)>& after_first_sort, int* launches)
{
    const int threads = 256, blocks = (n_max + threads - 1) / threads;
    const int32_t* cur_idx = in_idx; const int32_t* cur_cnt = in_count;
    int32_t* nxt_idx = list_a; int32_t* nxt_cnt = cnt_a;
    cudaError_t e;
    for (int pass = 0; pass < TB_PASSES; ++pass) {
        tband_reset_kernel<<<(TB_BINS + 255) / 256, 256, 0, st>>>(a, nxt_cnt);
        tband_key_kernel<<<blocks, threads, 0, st>>>(a, pass, cur_idx, cur_cnt);
        tband_scan_kernel<<<1, 1024, 0, st>>>(a);
        tband_scatter_kernel<<<blocks, threads, 0, st>>>(a, cur_idx, cur_cnt);
        *launches += 4;
        if ((e = cudaEventRecord(ev[TB_INST], st)) != cudaSuccess) return e;
        if (pass == 0 && after_first_sort && (e = after_first_sort(ev[TB_INST])) != cudaSuccess) return e;
        a.next_idx = nxt_idx; a.next_count = nxt_cnt;
        for (int inst = TB_INST - 1; inst >= 0; --inst) {
            TbandArgs x = a;
            x.scratch = a.scratch + plan.scratch_off[inst];
            x.scratch_stride = plan.stride[inst]; x.dir_bytes = plan.dir_bytes[inst];
            if ((e = cudaStreamWaitEvent(side[inst], ev[TB_INST], 0)) != cudaSuccess) return e;
            tband_kernel<<<plan.blocks[inst], TBAND_WARPS * 32, plan.smem[inst], side[inst]>>>(x, inst, tb_instance_nb(inst), pass);
            if ((e = cudaEventRecord(ev[inst], side[inst])) != cudaSuccess) return e;
            *launches += 1;
        }
        for (int inst = 0; inst < TB_INST; ++inst)
            if ((e = cudaStreamWaitEvent(st, ev[inst], 0)) != cudaSuccess) return e;
        cur_idx = nxt_idx; cur_cnt = nxt_cnt;
        if (nxt_idx == list_a) { nxt_idx = list_b; nxt_cnt = cnt_b; } else { nxt_idx = list_a; nxt_cnt = cnt_a; }
    }
    return cudaGetLastError();
}

}  // namespace sswb
