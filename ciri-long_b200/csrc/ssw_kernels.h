// ssw_kernels.h -- launch interfaces between the host orchestration (ssw_api.cu) and the kernels.
#pragma once
#include <functional>
#include "ssw_common.cuh"

namespace sswb {

// ---- score passes (ssw_score_impl.cuh, instances in ssw_score_[a-d].cu)
// ---- long references: the forward pass over column chunks
// A zero-started pass over columns [c0 - ov, c1) gives the exact H values of columns [c0, c1) when
// ov > m * (1 + maxMatch / gap_extend): a path that starts before the overlap has crossed more than ov
// columns, which costs at least (ov - m) * gap_extend in horizontal gaps against at most m * maxMatch gained,
// so its score is negative and it cannot set any H (floor 0).  Every chunk is an independent task; the
// tasks of a pair merge their best cell with an atomic max and write their columns of the pair's
// column records, the last one to finish runs the pair's epilogue.
struct ChunkPlan {
    int32_t chunk_cols;            // 0 = whole pairs
    int32_t max_match;
    int32_t* task_pair;            // task tables of this launch
    int32_t* task_c0;
    int32_t* task_c1;
    unsigned long long* pair_key;  // per pair, forward: (score, first column, row) of the best cell so far;
                                   //           reverse: (first task << 32) | number of tasks
    int32_t* pair_left;            // per pair: tasks still running
    const int64_t* col_off;        // per pair: offset of its column records in col_pool
    unsigned* col_pool;
    int4* task_res;                // reverse pass: per task (stop column, best score, its column, its row)
};
__host__ __device__ inline int chunk_overlap(int m, int maxMatch, int ge) { return m * (1 + (maxMatch + ge - 1) / ge) + 2; }
// reverse pass: columns of the first, bounded look (the stop column is normally about one alignment away)
__host__ __device__ inline int rev_look(int m) { return (2 * m + 1024 + RP_CHUNK - 1) / RP_CHUNK * RP_CHUNK; }
// number of tasks of a pair; 1 = the whole pair (queries of more than one tile, short references, or an
// overlap that would eat the gain)
__host__ __device__ inline int chunk_tasks(int m, int n, int chunk_cols, int maxMatch, int ge)
{
    if (chunk_cols <= 0 || m > VSTRIPS * KMAX || n <= chunk_cols || n >= (1 << 28)) return 1;
    if (4LL * (chunk_overlap(m, maxMatch, ge) + RP_CHUNK) > chunk_cols) return 1;
    return (n + chunk_cols - 1) / chunk_cols;
}

struct ScoreArgs {
    BatchView b;
    Scoring sc;
    WorkList wl;
    unsigned char* scratch;        // per-warp scratch, (blocks * SCORE_WARPS) * scratch_stride bytes
    long long scratch_stride;
    long long off_col, off_bnd, off_snap;   // byte offsets of the sub-buffers inside one warp's scratch
    int32_t rerun;                 // GOTOH forward only: 1 = deciding pass for a PS_NEED_GOTOH pair
    int32_t* next_idx;             // TRUNC forward only: list of pairs that need the deciding GOTOH pass
    const int32_t* next_base;      //   (same partitioning as the forward lists)
    int32_t* next_count;
    int32_t* wide_idx;             // forward only: pairs handed to the 32-bit kernels (ssw_score32.cu)
    int32_t* wide_count;
    ChunkPlan ck;                  // forward only: long references cut into column chunks (chunk_cols > 0)
};

// bytes of scratch one warp needs for references of up to n_cap columns
inline long long score_scratch_layout(int n_cap, long long* off_col, long long* off_bnd, long long* off_snap)
{
    long long o = 0;
    *off_col = o; o += (long long)n_cap * 4;               // column records (colmax | H last row)
    o = (o + 15) & ~15LL;
    *off_bnd = o; o += (long long)n_cap * 8;               // tile boundary (H, F, colmax)
    *off_snap = o; o += 2LL * KMAX * 32 * 4;               // best-column snapshots
    return (o + 127) & ~127LL;
}

cudaError_t launch_score(int K, bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st);

cudaError_t launch_score32(bool trunc, bool rev, const ScoreArgs& a, int blocks, cudaStream_t st);

// Which recurrence runs first on a pair when gap_open == gap_extend (0 = GOTOH, 1 = TRUNC).  Either order is
// exact: a GOTOH pass that overflows 8 bits hands the pair to TRUNC, a TRUNC pass that stays below hands it
// to the deciding GOTOH pass.  The guess only avoids the second pass: queries whose perfect-match score is
// at least 4/3 of the 8-bit limit are expected to overflow (noisy long-read alignments score ~0.7 per base).
__host__ __device__ inline int first_pass_kind(int m, int go, int ge, int maxScore, int bias)
{
    return (go == ge && 3LL * m * maxScore >= 4LL * (255 - bias)) ? 1 : 0;
}

// strip height (template parameter K of the score kernel) for a query of m rows; must match ssw_score_impl.cuh
__host__ __device__ inline int strip_height_for(int m, int trunc)
{
    if (!trunc) return m <= VSTRIPS * KMAX ? (m + VSTRIPS - 1) / VSTRIPS : KMAX;
    const int segLen = (m + 7) / 8;
    const int G = segLen > 8 * KMAX ? (segLen + KMAX - 1) / KMAX : 8;
    return (segLen + G - 1) / G;
}

// ---- work-list construction (ssw_lists.cu)
// list id = cls * 34 + kind * 17 + K   (cls: 0 normal / 1 long reference; kind: 0 GOTOH / 1 TRUNC; K: 1..16)
constexpr int N_LISTS = 2 * 2 * 17 + 5;           // + 68/69: reverse lists of the 32-bit kernels; 70/71: their forward counters; 72: tiny pairs
constexpr int LIST_WIDE32 = 68;
constexpr int LIST_TINY = 72;
// ---- tiny pairs: one pair per thread, forward + reverse fused (ssw_tiny.cu)
constexpr int TINY_LEN = 64;                      // both sequences at most this long ...
__host__ __device__ inline bool is_tiny_pair(int m, int n, int maxScore, int bias)
{   // ... and a score that cannot leave the 8-bit range: the byte flavour answers (ssw.c:805-809)
    return m >= 1 && n >= 1 && m <= TINY_LEN && n <= TINY_LEN && (long long)(m < n ? m : n) * maxScore + bias < 255;
}
struct TinyArgs {
    BatchView b;
    Scoring sc;
    WorkList wl;
};
cudaError_t launch_tiny(const TinyArgs& a, int blocks, cudaStream_t st);
__host__ __device__ inline int list_id(int cls, int kind, int K) { return cls * 34 + kind * 17 + K; }

struct ListSet {
    int32_t* idx;        // n_pairs entries, partitioned by list
    int32_t* count;      // N_LISTS
    int32_t* base;       // N_LISTS
    int32_t* fill;       // N_LISTS (scatter cursors)
    int32_t* cursor;     // N_LISTS (work-fetch cursors, zeroed by build)
};

// stage 0 = forward lists from (q_len, scoring); stage 1 = reverse lists from the forward results
cudaError_t build_lists(int stage, const BatchView& b, const Scoring& sc, int long_ref_threshold,
                        const ListSet& ls, cudaStream_t st, int* launches);
// all pairs that still need the CIGAR pass, in one list (count at ls.count[0])
// long references: expand the pairs of a forward list into column-chunk tasks
cudaError_t expand_tasks(const WorkList& wl, int max_pairs, bool rev, const BatchView& b, const Scoring& sc, const ChunkPlan& ck,
                         int32_t* task_count, cudaStream_t st, int* launches);
cudaError_t build_band_list(const BatchView& b, const Scoring& sc, const ListSet& ls, cudaStream_t st, int* launches);

// ASCII letters -> codes 0..4, in place (ssw_wrap.py:234-252 on the device)
cudaError_t encode_ascii(int8_t* seqs, long long n, cudaStream_t st);

// 4-bit packed bases (two per byte, low nibble first) -> one code per byte (nibbles above 4 -> N)
cudaError_t unpack4(const unsigned char* packed, int8_t* codes, long long n_bytes, cudaStream_t st);

cudaError_t snapshot_count(int32_t* snap, const int32_t* count, cudaStream_t st);
cudaError_t rewind_cursor(int32_t* cursor, const int32_t* snap, cudaStream_t st);

// clear status bits (and the CIGAR window) of every pair, e.g. before the CIGAR pass is repeated
cudaError_t clear_status_bits(const BatchView& b, int bits, cudaStream_t st);

// ---- banded DP + traceback (ssw_band.cu)
struct BandArgs {
    BatchView b;
    Scoring sc;
    WorkList wl;
    unsigned char* scratch;        // per-warp direction matrix + cigar staging
    long long scratch_stride;
    long long dir_bytes;           // bytes of the direction matrix area inside one warp's scratch
    int32_t cigar_stage_cap;       // ops
    uint32_t* cigar_buf;           // device output buffer
    long long cigar_cap;
    unsigned long long* cigar_used;
    int32_t* next_idx;             // class 0: pairs handed to class 1 (bands of 129-256 diagonals)
    int32_t* next_count;
    int32_t* next2_idx;            // class 0 and 1: pairs handed to class 2
    int32_t* next2_count;
};
constexpr int BAND_WARPS = 8;
// cls 0: bands up to 128 diagonals; 1: up to 256; 2: everything the others handed over
cudaError_t launch_band(int cls, const BandArgs& a, int blocks, cudaStream_t st);

// ---- throughput CIGAR pass: one pair per lane (ssw_tband.cu, ssw_tband_core.h)
struct TbandArgs {
    BatchView b;
    Scoring sc;
    int32_t* keys;                 // per list entry: bin (or -1 = handed over)
    int32_t* sorted;               // list sorted by (band blocks, row pairs), largest first
    int32_t* bin_count;            // TB_BINS
    int32_t* bin_base;             // TB_BINS
    int32_t* seg;                  // per instance: base[8] | count[8] | cursor[8] | handed-over flag[8]
    int32_t min_pairs_scale;       // 16 = default thresholds for "too few pairs for a launch"; 0 = always launch
    int32_t* next_idx;             // pairs whose band doubles: list of the next pass
    int32_t* next_count;
    int32_t* fallback_idx;         // pairs handed to ssw_band.cu (class 2 instance: any band width)
    int32_t* fallback_count;
    int32_t* fallback1_idx;        // pairs handed to ssw_band.cu (class 1 instance: up to 256 diagonals)
    int32_t* fallback1_count;
    unsigned char* scratch;        // per-warp direction words + op staging
    long long scratch_stride;
    long long dir_bytes;
    int32_t row_pairs_cap;         // row pairs the direction scratch of one lane holds
    int32_t stage_cap;             // ops
    int32_t one;                   // the constant 1 as a run-time value (keeps half of the direction-bit adds on the FMA pipe)
    uint32_t* cigar_buf;
    long long cigar_cap;
    unsigned long long* cigar_used;
};
struct TbandPlan {
    int32_t row_pairs_cap, stage_cap;
    int32_t blocks[8], smem[8];
    long long dir_bytes[8], stride[8], scratch_off[8];
    long long scratch_bytes;
};
constexpr int TBAND_BINS = 32 * 256;
cudaError_t tband_plan(int device, int sms, int max_q, long long budget, TbandPlan* plan);
cudaError_t tband_configure();
constexpr int TBAND_INSTANCES = 5;                 // side streams a batch needs: one per instance (+1 for the hand-overs)
cudaError_t launch_tband(TbandArgs a, const TbandPlan& plan, const int32_t* in_idx, const int32_t* in_count, int n_max,
                         int32_t* list_a, int32_t* list_b, int32_t* cnt_a, int32_t* cnt_b, cudaStream_t st,
                         cudaStream_t* side, cudaEvent_t* ev, const std::function<cudaError_t(cudaEvent_t)>& after_first_sort, int* launches);

// ---- batched edit distance (edit_distance.cu)
constexpr int ED_MAXSYM = 16;          // distinct symbols per batch (4-bit codes)
constexpr int EDIT_THREADS = 128;
struct EditArgs {
    const unsigned char* seqs;     // raw bytes, indexed with the caller's offsets
    const int64_t* x_off;
    const int32_t* x_len;
    const int64_t* y_off;
    const int32_t* y_len;
    int32_t* out;
    const int32_t* idx;            // pairs of this launch
    int32_t count;
    signed char* carry;            // tiled instance: one horizontal delta per text column and warp
    long long carry_stride;
    unsigned char code[256];       // byte -> dense symbol code
};
// kind 0/1: one thread per pair, pattern <= 32 / <= 64; kind 2..5: 4/8/16/32 lanes per pair
cudaError_t launch_edit(int kind, const EditArgs& a, int blocks, cudaStream_t st);

// ---- DPX issue-rate probe (ssw_peak.cu)
cudaError_t dpx_peak_probe(double* lane_instr_per_s, cudaStream_t st);

}  // namespace sswb
