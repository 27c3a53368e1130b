// ssw_api.cu -- host side of libssw_cuda.so: the C ABI of include/ssw_cuda.h.
//
// Batched interface: ssw_batch_create uploads the struct-of-arrays batch, ssw_batch_run enqueues the
// whole hot path on one stream without any host round trip (work lists, counts and cursors live in
// device memory), ssw_batch_fetch synchronises and copies results + CIGARs back.
// Legacy interface: the six symbols of the reference libssw.so (ssw.h:72-182) as a batch of one pair.
// There is no CPU implementation of the alignment in this file or behind it.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ssw_cuda.h"
#include "ssw_common.cuh"
#include "ssw_kernels.h"

using namespace sswb;

static thread_local std::string g_last_error;
// SSW_CUDA_TRACE=1: print host-side phase timings of the batch calls to stderr (diagnostics)
static bool trace_on() { static int v = -1; if (v < 0) { const char* e = getenv("SSW_CUDA_TRACE"); v = (e && *e == '1') ? 1 : 0; } return v == 1; }
struct TraceTimer {
    const char* what; std::chrono::steady_clock::time_point t0;
    explicit TraceTimer(const char* w) : what(w), t0(std::chrono::steady_clock::now()) {}
    ~TraceTimer() { if (trace_on()) fprintf(stderr, "[ssw_cuda] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};
static void set_error(const std::string& s) { g_last_error = s; }
extern "C" const char* ssw_cuda_last_error(void) { return g_last_error.c_str(); }

#define CU_TRY(expr)                                                                          \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
            return SSW_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

// Device memory comes from a stream-ordered pool that belongs to this library (one per device, created on first
// use, thread-safe): repeated batches (the one-shot call creates one per chunk) get their buffers back from the
// pool instead of paying cudaMalloc, and the process-wide default pool of other CUDA users is left alone.
// ssw_cuda_trim_pools() hands the cached memory back to the driver.
static cudaMemPool_t g_pool[64];
static std::once_flag g_pool_once[64];
static cudaMemPool_t lib_pool(int dev)
{
    if (dev < 0 || dev >= 64) return nullptr;
    std::call_once(g_pool_once[dev], [dev]() {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
            unsigned long long thr = ~0ULL;                       // keep freed blocks cached for the next batch
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            g_pool[dev] = pool;
        } else { cudaGetLastError(); g_pool[dev] = nullptr; }
    });
    return g_pool[dev];
}
static cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t st)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = lib_pool(dev);
    if (pool) return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, st);
    return cudaMallocAsync(p, bytes ? bytes : 1, st);
}
extern "C" int ssw_cuda_trim_pools(void)
{
    for (int dev = 0; dev < 64; ++dev) if (g_pool[dev]) cudaMemPoolTrimTo(g_pool[dev], 0);
    return SSW_OK;
}
template <typename T> static cudaError_t dev_alloc_t(T** p, size_t count, cudaStream_t st) { return dev_alloc((void**)p, count * sizeof(T), st); }
static void dev_free(void* p, cudaStream_t st) { if (p) cudaFreeAsync(p, st); }

static const int LONG_REF_THRESHOLD = 32768;          // references longer than this get their own scratch class
static const long long SCRATCH_BUDGET = 3LL << 30;    // bytes of score-pass scratch per class

struct ssw_batch {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int32_t n = 0;
    int sms = 0;
    Scoring sc;
    // inputs (d_seqs holds bytes [seq_lo, seq_hi) of the caller's buffer; kernels index it through
    // d_seqs - seq_lo so the caller's offsets are used unchanged)
    long long seq_lo = 0, seq_hi = 0;
    std::vector<int32_t> h_mask;
    bool ascii_done = false;
    bool packed = false;                 // the caller's sequence buffer holds two bases per byte (offsets count bases)
    unsigned char* d_packed = nullptr;
    std::vector<int64_t> h_col_off;
    int8_t* d_seqs = nullptr;
    long long *d_qoff = nullptr, *d_roff = nullptr;
    int32_t *d_qlen = nullptr, *d_rlen = nullptr, *d_mask = nullptr;
    PairRec* d_rec = nullptr;
    // lists: [count | base | fill | cursor | count2 | cursor2] x N_LISTS
    int32_t *d_idx = nullptr, *d_idx2 = nullptr, *d_idx3 = nullptr, *d_idx4 = nullptr, *d_meta = nullptr;
    // scratch
    unsigned char* d_sscr[2] = {nullptr, nullptr};
    // the per-list launches of a stage run side by side (FAN streams): every stream has its own class-0 scratch; the
    // class-1 scratch (megabytes per warp) is cut into two halves instead
    static constexpr int FAN = 4;
    unsigned char* d_sscr0x[FAN] = {nullptr, nullptr, nullptr, nullptr};
    int fan0 = 1, fan1 = 1;
    cudaStream_t sside[FAN] = {};                    // (aliases of the first FAN entries of tside: the device has few hardware queues)
    cudaEvent_t sev[FAN + 1] = {};
    long long sstride[2] = {0, 0}, off_col[2] = {0, 0}, off_bnd[2] = {0, 0}, off_snap[2] = {0, 0};
    int sblocks[2] = {0, 0};
    unsigned char* d_bscr = nullptr;                 // CIGAR pass, narrow instance (bands up to 128 diagonals)
    long long bstride = 0, bdir = 0;
    int bblocks = 0, bstage = 0;
    unsigned char* d_wscr = nullptr;                 // CIGAR pass, wide instance (few pairs, big direction matrices)
    long long wstride = 0, wdir = 0;
    int wblocks = 0;
    // CIGAR pass, throughput instance (ssw_tband.cu): one pair per lane, used for large batches
    bool use_tband = false;
    bool no_cigar = false;                           // (flag & 4) with filterd < 0: no pair can pass the CIGAR gate (ssw.c:850)
    TbandPlan tplan;
    unsigned char* d_tscr = nullptr;
    cudaStream_t tside[TBAND_INSTANCES + 1] = {};    // one stream per instance + one for the hand-overs
    cudaEvent_t tev[TBAND_INSTANCES + 3] = {};
    int32_t *d_tlists = nullptr, *d_tbins = nullptr; // keys | sorted | list a | list b ;  bin_count | bin_base | seg | counts
    uint32_t* d_cigar = nullptr;
    long long cigar_cap = 0, cigar_worst = 0;
    unsigned long long* d_cigar_used = nullptr;
    // host-side shape summary
    int max_q = 0, max_r = 0, maxK = 0;
    int max_rows = 0;                    // bound on the rows of any pair's trimmed rectangle (CIGAR scratch)
    bool have[2][2][KMAX + 1];           // forward lists known to be non-empty: [cls][kind][K]
    bool have_t2[2][KMAX + 1];
    int32_t n_tiny = 0;                  // pairs of the one-pair-per-thread score kernel (ssw_tiny.cu)           // GOTOH-first pairs that may overflow and move on to TRUNC: [cls][K]
    // long references (class 1): forward pass over column chunks (ssw_kernels.h: ChunkPlan)
    int32_t chunk_cols = 0;
    int32_t long_pairs[2][KMAX + 1];     // class-1 pairs per forward list [kind][K]
    long long task_cap[2][KMAX + 1];     // and the tasks they expand to
    long long task_base[2][KMAX + 1];    // region of each list inside the task tables
    int64_t* d_col_off = nullptr;
    unsigned* d_col_pool = nullptr;
    unsigned long long* d_pair_key = nullptr;
    int32_t* d_pair_left = nullptr;
    int32_t* d_task = nullptr;           // task_pair | task_c0 | task_c1, each task_total entries
    long long task_total = 0;
    int32_t* d_task_meta = nullptr;      // counts[2][KMAX+1] | cursors[2][KMAX+1]
    int32_t* d_rtask = nullptr;          // reverse pass: task tables (one list at a time) and per-task results
    int4* d_rres = nullptr;
    long long rev_task_total = 0;
    int32_t long_total = 0;              // class-1 pairs in the batch
    int64_t launches = 0;
    PairRec* h_rec = nullptr;            // the caller's result array during a fetch: ssw_result and PairRec share their layout
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // stage boundaries of the last run

    BatchView view() const { return BatchView{d_seqs - seq_lo, d_qoff, d_qlen, d_roff, d_rlen, d_mask, d_rec, n}; }
    ListSet lists() const { return ListSet{d_idx, d_meta, d_meta + N_LISTS, d_meta + 2 * N_LISTS, d_meta + 3 * N_LISTS}; }
    int32_t* count2() const { return d_meta + 4 * N_LISTS; }
    int32_t* cursor2() const { return d_meta + 5 * N_LISTS; }
    int32_t* count3() const { return d_meta + 6 * N_LISTS; }
    int32_t* cursor3() const { return d_meta + 7 * N_LISTS; }
};


// Which scoring schemes the device recurrences reproduce bit-exactly (see ssw_score_impl.cuh header):
// 5x5 matrix with a zero N row/column (the only matrix ssw_wrap.py:146-159 builds), gap_open >=
// gap_extend >= 1 and 2*gap_extend >= the largest mismatch penalty (an insertion next to a deletion is
// then never better than substitutions, which makes the reference's lazy-F bookkeeping unobservable).
static bool scoring_supported(const ssw_scoring* s, std::string* why)
{
    int minv = 0;
    for (int k = 0; k < 25; ++k) minv = std::min<int>(minv, s->mat[k]);
    for (int k = 0; k < 5; ++k)
        if (s->mat[4 * 5 + k] != 0 || s->mat[k * 5 + 4] != 0) { *why = "N row/column of the matrix must be zero"; return false; }
    if (s->gap_extend < 1 || s->gap_open < s->gap_extend) { *why = "need gap_open >= gap_extend >= 1"; return false; }
    if (2 * (int)s->gap_extend < -minv) { *why = "need 2*gap_extend >= largest mismatch penalty"; return false; }
    if (s->gap_open > 100 || -minv > 100) { *why = "penalties above 100 are not supported"; return false; }
    return true;
}

extern "C" void ssw_batch_destroy(ssw_batch* b)
{
    if (!b) return;
    TraceTimer tt("batch_destroy");
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaStream_t fs = b->stream;
    dev_free(b->d_seqs, fs); dev_free(b->d_qoff, fs); dev_free(b->d_roff, fs); dev_free(b->d_qlen, fs); dev_free(b->d_rlen, fs);
    dev_free(b->d_mask, fs); dev_free(b->d_rec, fs); dev_free(b->d_idx, fs); dev_free(b->d_idx2, fs); dev_free(b->d_idx3, fs); dev_free(b->d_idx4, fs); dev_free(b->d_meta, fs);
    dev_free(b->d_col_off, fs); dev_free(b->d_col_pool, fs); dev_free(b->d_pair_key, fs); dev_free(b->d_pair_left, fs);
    dev_free(b->d_task, fs); dev_free(b->d_task_meta, fs); dev_free(b->d_rtask, fs); dev_free(b->d_rres, fs);
    dev_free(b->d_packed, fs); dev_free(b->d_tscr, fs); dev_free(b->d_tlists, fs); dev_free(b->d_tbins, fs);
    dev_free(b->d_sscr[0], fs); dev_free(b->d_sscr[1], fs); dev_free(b->d_bscr, fs); dev_free(b->d_wscr, fs); dev_free(b->d_cigar, fs); dev_free(b->d_cigar_used, fs);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (int k = 0; k < 5; ++k) if (b->ev[k]) cudaEventDestroy(b->ev[k]);
    for (int k = 1; k < ssw_batch::FAN; ++k) dev_free(b->d_sscr0x[k], fs);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (auto& x : b->sev) if (x) cudaEventDestroy(x);
    for (auto& x : b->tside) if (x) cudaStreamDestroy(x);
    for (auto& x : b->tev) if (x) cudaEventDestroy(x);
    if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

static int batch_alloc(ssw_batch* b, const int8_t* seqs, int64_t seqs_len, const int64_t* q_off, const int32_t* q_len,
                       const int64_t* r_off, const int32_t* r_len, const int32_t* mask_len)
{
    const int32_t n = b->n;
    TraceTimer tt("batch_alloc (validate+alloc+h2d enqueue)");
    b->h_mask.resize(n);
    memset(b->have, 0, sizeof b->have);
    memset(b->have_t2, 0, sizeof b->have_t2);
    int maxScore = 0;
    for (int k = 0; k < 25; ++k) maxScore = std::max<int>(maxScore, b->sc.mat[k]);
    long long cig_worst = 16, q_total = 0;
    long long lo = seqs_len, hi = 0;
    long long long_cols = 0;
    memset(b->long_pairs, 0, sizeof b->long_pairs);
    for (int32_t p = 0; p < n; ++p) if (q_len[p] > 0 && r_len[p] > LONG_REF_THRESHOLD) long_cols += r_len[p];
    if (long_cols > 0) {
        // enough tasks to fill the machine a few times over, chunks long enough to amortise the overlap
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device);
        // (a launch cannot end before its longest task does, and every strip-height list is its own launch: 16 k
        // columns keep that tail at a few ms while the warm-up overlap of a task stays below ~10 %)
        long long c = long_cols / (8LL * std::max(sms, 1) * SCORE_WARPS);
        c = std::max<long long>(8192, std::min<long long>(c, 16384));
        b->chunk_cols = (int32_t)((c + RP_CHUNK - 1) / RP_CHUNK * RP_CHUNK);
    }
    memset(b->task_cap, 0, sizeof b->task_cap);
    std::vector<int64_t> h_col_off;
    long long col_total = 0;
    if (b->chunk_cols) h_col_off.assign(n, 0);
    // Shape summary of the batch (which kernel instances it needs, scratch bounds, byte range to upload).  One pass
    // over the pairs, on several host threads for large batches: with millions of tiny pairs this loop, not the
    // GPU, would otherwise set the pace of the one-shot call.
    struct Part {
        bool have[2][2][KMAX + 1] = {}; bool have_t2[2][KMAX + 1] = {};
        int32_t long_pairs[2][KMAX + 1] = {}; long long task_cap[2][KMAX + 1] = {};
        int32_t n_tiny = 0, long_total = 0, bad = -1;
        int max_q = 0, max_r = 0, max_rows = 0, maxK = 0;
        long long cig_worst = 0, q_total = 0, rev_task_total = 0, lo = 0, hi = 0;
    };
    const Scoring sc = b->sc;
    const int32_t chunk_cols = b->chunk_cols;
    int32_t* h_mask = b->h_mask.data();
    auto scan = [&](int32_t p0, int32_t p1, Part& P) {
        P.lo = seqs_len; P.hi = 0;
        for (int32_t p = p0; p < p1; ++p) {
            const int m = q_len[p], r = r_len[p];
            if (m < 0 || r < 0 || q_off[p] < 0 || r_off[p] < 0 || q_off[p] + m > seqs_len || r_off[p] + r > seqs_len) { if (P.bad < 0) P.bad = p; continue; }
            h_mask[p] = mask_len ? mask_len[p] : (m > 30 ? m / 2 : 15);      // ssw_wrap.py:196-199
            P.max_q = std::max(P.max_q, m);
            P.max_r = std::max(P.max_r, r);
            P.lo = std::min<long long>(P.lo, std::min(q_off[p], r_off[p]));
            P.hi = std::max<long long>(P.hi, std::max(q_off[p] + m, r_off[p] + r));
            if (is_tiny_pair(m, r, maxScore, sc.bias)) {
                P.n_tiny += 1;
                P.cig_worst += 2LL * m + 3;
                P.q_total += m;
                P.max_rows = std::max<int>(P.max_rows, m);
            } else if (m > 0 && r > 0) {
                const int kind = first_pass_kind(m, sc.go, sc.ge, maxScore, sc.bias);
                const int K = strip_height_for(m, kind);
                const int cls = r > LONG_REF_THRESHOLD ? 1 : 0;
                P.have[cls][kind][K] = true;
                if (cls) {
                    P.long_pairs[kind][K] += 1;
                    P.task_cap[kind][K] += chunk_tasks(m, r, chunk_cols, maxScore, sc.ge);
                    P.long_total += 1;
                    P.rev_task_total += (r + chunk_cols - 1) / chunk_cols + 1;
                }
                if (kind == 0 && sc.go == sc.ge && (long long)m * maxScore + sc.bias >= 255) P.have_t2[cls][strip_height_for(m, 1)] = true;
                P.maxK = std::max(P.maxK, std::max(K, strip_height_for(m, 0)));
                P.cig_worst += 2LL * m + 3;
                P.q_total += m;
                // rows of the trimmed rectangle = aligned query span <= query length, and a local alignment cannot hold
                // more inserted query bases than its matches pay for: span <= r * (1 + maxScore / gap_extend)
                const long long span = std::min<long long>(m, (long long)r * (1 + (maxScore + sc.ge - 1) / std::max(1, (int)sc.ge)) + 1);
                P.max_rows = std::max<int>(P.max_rows, (int)span);
            }
        }
    };
    int nthreads = 1;
    if (n >= 262144) nthreads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    std::vector<Part> parts(nthreads);
    if (nthreads == 1) scan(0, n, parts[0]);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back(scan, (int32_t)((long long)n * t / nthreads), (int32_t)((long long)n * (t + 1) / nthreads), std::ref(parts[t]));
        for (auto& x : th) x.join();
    }
    for (const Part& P : parts) {
        if (P.bad >= 0) { set_error("pair " + std::to_string(P.bad) + ": offsets/lengths outside the sequence buffer"); return SSW_ERR_ARG; }
        for (int c = 0; c < 2; ++c) for (int K = 0; K <= KMAX; ++K) {
            b->have_t2[c][K] |= P.have_t2[c][K];
            b->long_pairs[c][K] += P.long_pairs[c][K]; b->task_cap[c][K] += P.task_cap[c][K];
            for (int k = 0; k < 2; ++k) b->have[c][k][K] |= P.have[c][k][K];
        }
        b->n_tiny += P.n_tiny; b->long_total += P.long_total; b->rev_task_total += P.rev_task_total;
        b->max_q = std::max(b->max_q, P.max_q); b->max_r = std::max(b->max_r, P.max_r);
        b->max_rows = std::max(b->max_rows, P.max_rows); b->maxK = std::max(b->maxK, P.maxK);
        cig_worst += P.cig_worst; q_total += P.q_total;
        lo = std::min(lo, P.lo); hi = std::max(hi, P.hi);
    }
    if (b->chunk_cols)                                   // column records of the long references: offsets in pair order
        for (int32_t p = 0; p < n; ++p) {
            const int m = q_len[p], r = r_len[p];
            if (m > 0 && r > LONG_REF_THRESHOLD && !is_tiny_pair(m, r, maxScore, sc.bias)) {
                h_col_off[p] = col_total;
                col_total += ((long long)r + 31) & ~31LL;             // 128-byte aligned: no cache line is shared by two pairs
            }
        }
    if (hi < lo) { lo = 0; hi = 0; }
    if (b->packed) lo &= ~31LL;                          // packed upload: start on a byte (and 16-byte) boundary of the packed buffer
    b->seq_lo = lo; b->seq_hi = hi;
    // Allocation sizes are functions of these three numbers; they are rounded up so that consecutive
    // batches of similar shape request identical blocks and the stream-ordered pool can hand the same
    // memory back instead of growing.
    const int cap_q = (b->max_q + 255) & ~255, cap_r = (b->max_r + 255) & ~255;
    const size_t cap_seq = ((size_t)(hi - lo) + 64 + (4u << 20)) & ~(size_t)((4u << 20) - 1);
    CU_TRY(cudaDeviceGetAttribute(&b->sms, cudaDevAttrMultiProcessorCount, b->device));

    cudaStream_t st = b->stream;
    const size_t nn = ((size_t)std::max(n, 1) + 16383) & ~(size_t)16383;
    CU_TRY(dev_alloc_t(&b->d_qoff, nn, st)); CU_TRY(dev_alloc_t(&b->d_roff, nn, st));
    CU_TRY(dev_alloc_t(&b->d_qlen, nn, st)); CU_TRY(dev_alloc_t(&b->d_rlen, nn, st)); CU_TRY(dev_alloc_t(&b->d_mask, nn, st));
    CU_TRY(dev_alloc_t(&b->d_rec, nn, st));
    CU_TRY(dev_alloc_t(&b->d_idx, nn, st)); CU_TRY(dev_alloc_t(&b->d_idx2, nn, st)); CU_TRY(dev_alloc_t(&b->d_idx3, 2 * nn, st));
    CU_TRY(dev_alloc_t(&b->d_meta, 8 * N_LISTS, st)); CU_TRY(dev_alloc_t(&b->d_idx4, nn, st));
    CU_TRY(dev_alloc_t(&b->d_cigar_used, 1, st));
    CU_TRY(dev_alloc_t(&b->d_seqs, cap_seq, st));
    CU_TRY(cudaMemcpyAsync(b->d_qoff, q_off, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(b->d_roff, r_off, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(b->d_qlen, q_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(b->d_rlen, r_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(b->d_mask, b->h_mask.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    // only the bytes the pairs of this batch reference travel; with pinned caller memory this copy is
    // asynchronous (the caller keeps `seqs` alive until ssw_batch_fetch, like every CUDA async copy)
    if (hi > lo && !b->packed) CU_TRY(cudaMemcpyAsync(b->d_seqs, seqs + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, st));
    if (hi > lo && b->packed) {
        // half the bytes cross PCIe; the device expands them to one code per byte (what every kernel reads)
        const long long nb = (hi - lo + 1) / 2;
        CU_TRY(dev_alloc_t(&b->d_packed, (size_t)nb + 16, st));
        CU_TRY(cudaMemcpyAsync(b->d_packed, reinterpret_cast<const unsigned char*>(seqs) + lo / 2, (size_t)nb, cudaMemcpyHostToDevice, st));
        CU_TRY(unpack4(b->d_packed, b->d_seqs, nb, st));
    }

    if (b->chunk_cols) {
        b->task_total = 0;
        for (int kind = 0; kind < 2; ++kind)
            for (int K = 1; K <= KMAX; ++K) { b->task_base[kind][K] = b->task_total; b->task_total += b->task_cap[kind][K]; }
        CU_TRY(dev_alloc_t(&b->d_col_off, nn, st));
        CU_TRY(dev_alloc_t(&b->d_col_pool, (size_t)col_total + 64, st));
        CU_TRY(dev_alloc_t(&b->d_pair_key, nn, st));
        CU_TRY(dev_alloc_t(&b->d_pair_left, nn, st));
        CU_TRY(dev_alloc_t(&b->d_task, (size_t)(3 * b->task_total + 16), st));
        CU_TRY(dev_alloc_t(&b->d_task_meta, 4 * (KMAX + 1), st));
        if (b->sc.flag != 0) {
            CU_TRY(dev_alloc_t(&b->d_rtask, (size_t)(3 * b->rev_task_total + 16), st));
            CU_TRY(dev_alloc_t(&b->d_rres, (size_t)(b->rev_task_total + 16), st));
        }
        // (the vector dies with this function: the copy is made from a staging allocation that outlives it)
        b->h_col_off.swap(h_col_off);
        CU_TRY(cudaMemcpyAsync(b->d_col_off, b->h_col_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    }
    // score-pass scratch: class 0 = references up to LONG_REF_THRESHOLD columns, class 1 = longer ones
    for (int cls = 0; cls < 2; ++cls) {
        bool any = false;
        for (int kind = 0; kind < 2; ++kind) for (int K = 1; K <= KMAX; ++K) any |= b->have[cls][kind][K];
        if (!any) continue;
        const int ncap = cls == 0 ? std::min(cap_r, LONG_REF_THRESHOLD) : cap_r;
        b->sstride[cls] = score_scratch_layout(ncap, &b->off_col[cls], &b->off_bnd[cls], &b->off_snap[cls]);
        // long references need megabytes of column records per resident warp: let them take up to a quarter
        // of the free HBM, so that every SM keeps its full set of warps
        long long budget = SCRATCH_BUDGET;
        if (cls == 1) {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
                budget = std::max<long long>(budget, std::min<long long>((long long)(free_b / 4), 40LL << 30));
        }
        long long blocks = budget / (b->sstride[cls] * SCORE_WARPS);
        blocks = std::max<long long>(1, std::min<long long>(blocks, b->sms));
        blocks = std::min<long long>(blocks, (n + 15) / 16);
        b->sblocks[cls] = (int)blocks;
        const size_t bytes = (size_t)(blocks * SCORE_WARPS * b->sstride[cls]);
        CU_TRY(dev_alloc_t(&b->d_sscr[cls], bytes, st));
        // SSW_CUDA_FANOUT: 0 = off, 1 (default) = long references only, 2 = short references as well.  Measured: long
        // references (S1) 489 -> 402 ms per step; short ones gain 1-3 % in the score stages of C2 / C5 but the CIGAR
        // stage that follows ran 5-8 % slower on mixed batches for reasons not understood, so they stay on one stream.
        int fan_mode = n >= 4096 ? 1 : 0;                // small batches: one stream, nothing to overlap
        if (const char* e = getenv("SSW_CUDA_FANOUT")) fan_mode = std::min(fan_mode == 0 ? 0 : 2, atoi(e));
        const bool fan = fan_mode >= 1;
        if (cls == 0 && fan_mode >= 2 && bytes * ssw_batch::FAN <= (4ULL << 30)) {
            b->d_sscr0x[0] = b->d_sscr[0];
            for (int k = 1; k < ssw_batch::FAN; ++k) CU_TRY(dev_alloc_t(&b->d_sscr0x[k], bytes, st));
            b->fan0 = ssw_batch::FAN;
        }
        if (cls == 1 && fan && blocks >= 8) b->fan1 = 2;
    }
    // One set of side streams serves the score stages (first FAN of them) and the CIGAR stage (all): streams beyond
    // the device's hardware queues (8 by default) would share queues and serialise what is meant to overlap.
    if (b->fan0 > 1 || b->fan1 > 1) {
        for (auto& x : b->tside) if (!x) CU_TRY(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        for (int k = 0; k < ssw_batch::FAN; ++k) b->sside[k] = b->tside[k];
        for (auto& x : b->sev) CU_TRY(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    }
    // CIGAR stage scratch and output
    b->no_cigar = (b->sc.flag & 4) != 0 && b->sc.filterd < 0;              // begin coordinates only: nothing passes ssw.c:850
    if (b->sc.flag != 0 && !b->no_cigar) {
        b->bstage = cap_q + cap_r + 16;
        const long long rows = std::min(cap_q, cap_r + cap_q);           // rows of the trimmed rectangle <= query length
        // large batches: one pair per lane (ssw_tband.cu); small ones: one pair per warp (ssw_band.cu)
        long tband_min = 16384;
        if (const char* e = getenv("SSW_CUDA_TBAND_MIN")) tband_min = atol(e);
        b->use_tband = n >= tband_min;
        long long blocks = 0;
        if (b->use_tband) {
            long long budget = 12LL << 30;               // direction words of the resident lock-step rounds (HBM: 180 GB)
            if (const char* e = getenv("SSW_CUDA_TBAND_BUDGET_MB")) { const long v = atol(e); if (v > 0) budget = (long long)v << 20; }
            CU_TRY(tband_plan(b->device, b->sms, b->max_rows, budget, &b->tplan));
            CU_TRY(tband_configure());
            for (auto& x : b->tside) if (!x) CU_TRY(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
            for (auto& x : b->tev) CU_TRY(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
            CU_TRY(dev_alloc_t(&b->d_tscr, (size_t)b->tplan.scratch_bytes, st));
            CU_TRY(dev_alloc_t(&b->d_tlists, 4 * nn, st));
            CU_TRY(dev_alloc_t(&b->d_tbins, 2 * (size_t)TBAND_BINS + 64, st));
        } else {
            // narrow instance: row stride <= 128 bytes; wide instance: up to 1024 diagonals (wider bands, or
            // longer reads, are re-run from ssw_batch_fetch with a scratch sized for them)
            b->bdir = std::max<long long>(65536, 128LL * rows);
            b->bstride = ((long long)b->bstage * 4 + b->bdir + 255) & ~255LL;
            blocks = std::min<long long>((long long)b->sms * 4, (n + BAND_WARPS - 1) / BAND_WARPS);
            blocks = std::max<long long>(1, std::min<long long>(blocks, SCRATCH_BUDGET / (b->bstride * BAND_WARPS)));
            b->bblocks = (int)blocks;
            CU_TRY(dev_alloc_t(&b->d_bscr, (size_t)(blocks * BAND_WARPS * b->bstride), st));
        }
        b->wdir = std::min<long long>(1024LL * rows + 65536, 64LL << 20);
        b->wstride = ((long long)b->bstage * 4 + b->wdir + 255) & ~255LL;
        blocks = std::min<long long>((long long)b->sms * 2, (n + BAND_WARPS - 1) / BAND_WARPS);
        blocks = std::max<long long>(1, std::min<long long>(blocks, (b->use_tband ? 4LL << 30 : SCRATCH_BUDGET) / (b->wstride * BAND_WARPS)));
        b->wblocks = (int)blocks;
        CU_TRY(dev_alloc_t(&b->d_wscr, (size_t)(blocks * BAND_WARPS * b->wstride), st));
        // CIGAR output: the worst case is 2*len(query)+3 ops per pair; real alignments need a small
        // fraction of that, so start with an estimate and grow to the worst case only if a run overflows
        b->cigar_worst = cig_worst;
        b->cigar_cap = std::min<long long>(cig_worst, (std::max<long long>(1 << 16, q_total / 2 + 16LL * n) + (1 << 20)) & ~((1LL << 20) - 1));
        CU_TRY(dev_alloc_t(&b->d_cigar, (size_t)b->cigar_cap, st));
    }
    return SSW_OK;
}

static ssw_batch* batch_create_impl(int device, void* stream, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                                    const int64_t* q_off, const int32_t* q_len, const int64_t* r_off,
                                    const int32_t* r_len, const int32_t* mask_len, const ssw_scoring* scoring, bool packed);

extern "C" ssw_batch* ssw_batch_create(int device, void* stream, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                                       const int64_t* q_off, const int32_t* q_len, const int64_t* r_off,
                                       const int32_t* r_len, const int32_t* mask_len, const ssw_scoring* scoring)
{
    return batch_create_impl(device, stream, n_pairs, seqs, seqs_len, q_off, q_len, r_off, r_len, mask_len, scoring, false);
}

extern "C" ssw_batch* ssw_batch_create_packed(int device, void* stream, int32_t n_pairs, const uint8_t* packed, int64_t n_bases,
                                              const int64_t* q_off, const int32_t* q_len, const int64_t* r_off,
                                              const int32_t* r_len, const int32_t* mask_len, const ssw_scoring* scoring)
{
    return batch_create_impl(device, stream, n_pairs, reinterpret_cast<const int8_t*>(packed), n_bases, q_off, q_len, r_off, r_len,
                             mask_len, scoring, true);
}

extern "C" void ssw_pack_dna4(const int8_t* codes, int64_t n, uint8_t* packed)
{
    for (int64_t k = 0; k + 1 < n; k += 2) {
        const unsigned a = (unsigned char)codes[k] > 4 ? 4u : (unsigned char)codes[k], c = (unsigned char)codes[k + 1] > 4 ? 4u : (unsigned char)codes[k + 1];
        packed[k >> 1] = (uint8_t)(a | (c << 4));
    }
    if (n & 1) packed[n >> 1] = (uint8_t)((unsigned char)codes[n - 1] > 4 ? 4 : codes[n - 1]);
}

static ssw_batch* batch_create_impl(int device, void* stream, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                                    const int64_t* q_off, const int32_t* q_len, const int64_t* r_off,
                                    const int32_t* r_len, const int32_t* mask_len, const ssw_scoring* scoring, bool packed)
{
    if (n_pairs < 0 || !scoring || (n_pairs > 0 && (!seqs || !q_off || !q_len || !r_off || !r_len))) {
        set_error("ssw_batch_create: invalid argument");
        return nullptr;
    }
    std::string why;
    if (!scoring_supported(scoring, &why)) {
        set_error("scoring scheme not supported by the device kernels: " + why);
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        set_error("no usable CUDA device (libssw_cuda has no CPU path)");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    ssw_batch* b = new ssw_batch();
    b->device = device;
    b->n = n_pairs;
    b->packed = packed;
    memcpy(b->sc.mat, scoring->mat, 25);
    b->sc.go = scoring->gap_open; b->sc.ge = scoring->gap_extend;
    int minv = 0;
    for (int k = 0; k < 25; ++k) minv = std::min<int>(minv, scoring->mat[k]);
    b->sc.bias = -minv;                                                  // ssw.c:756-762
    b->sc.flag = scoring->flag; b->sc.filters = scoring->filters; b->sc.filterd = scoring->filterd;
    if (stream) b->stream = (cudaStream_t)stream;
    else {
        if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream"); delete b; return nullptr; }
        b->own_stream = true;
    }
    if (batch_alloc(b, seqs, seqs_len, q_off, q_len, r_off, r_len, mask_len) != SSW_OK) { ssw_batch_destroy(b); return nullptr; }
    for (int k = 0; k < 5; ++k)
        if (cudaEventCreate(&b->ev[k]) != cudaSuccess) { set_error("cudaEventCreate"); ssw_batch_destroy(b); return nullptr; }
    return b;
}

// CIGAR pass: list of pairs that pass the reference's gate (ssw.c:850), narrow instance, then the two
// wider instances for the pairs handed over.
static int enqueue_cigar_stage(ssw_batch* b, int* launches)
{
    cudaStream_t st = b->stream;
    const BatchView view = b->view();
    const ListSet ls = b->lists();
    CU_TRY(cudaMemsetAsync(b->d_cigar_used, 0, 8, st));
    CU_TRY(build_band_list(view, b->sc, ls, st, launches));
    BandArgs ba;
    ba.b = view; ba.sc = b->sc;
    ba.wl = WorkList{ls.idx, nullptr, ls.count, ls.cursor};
    ba.scratch = b->d_bscr; ba.scratch_stride = b->bstride; ba.dir_bytes = b->bdir;
    ba.cigar_stage_cap = b->bstage; ba.cigar_buf = b->d_cigar; ba.cigar_cap = b->cigar_cap;
    ba.cigar_used = b->d_cigar_used;
    // class 0 hands bands of 129-256 diagonals to class 1 (list in d_idx2) and anything wider to class 2
    // (list in d_idx4); class 1 hands on to class 2 as well
    ba.next_idx = b->d_idx2; ba.next_count = b->count2();
    ba.next2_idx = b->d_idx4; ba.next2_count = b->count3();
    CU_TRY(cudaMemsetAsync(b->count2(), 0, 4 * N_LISTS * 4, st));       // count2, cursor2, count3, cursor3
    if (b->use_tband) {
        // throughput instance first; what it hands over (bands wider than 126, scores near the 16-bit range,
        // score-0 pairs) lands in the class-2 list
        const size_t nn = ((size_t)std::max(b->n, 1) + 16383) & ~(size_t)16383;
        TbandArgs ta;
        ta.b = view; ta.sc = b->sc;
        ta.keys = b->d_tlists; ta.sorted = b->d_tlists + nn;
        ta.bin_count = b->d_tbins; ta.bin_base = b->d_tbins + TBAND_BINS; ta.seg = b->d_tbins + 2 * TBAND_BINS;
        ta.next_idx = nullptr; ta.next_count = nullptr;
        ta.fallback_idx = b->d_idx4; ta.fallback_count = b->count3();
        ta.fallback1_idx = b->d_idx2; ta.fallback1_count = b->count2();
        ta.min_pairs_scale = 16;
        if (const char* e = getenv("SSW_CUDA_TBAND_SEG_SCALE")) ta.min_pairs_scale = atoi(e);
        ta.scratch = b->d_tscr; ta.scratch_stride = 0; ta.dir_bytes = 0;
        ta.row_pairs_cap = b->tplan.row_pairs_cap; ta.stage_cap = b->tplan.stage_cap; ta.one = 1;
        ta.cigar_buf = b->d_cigar; ta.cigar_cap = b->cigar_cap; ta.cigar_used = b->d_cigar_used;
        int32_t* cnt = b->d_tbins + 2 * TBAND_BINS + 32;
        // the warp-per-pair instances start on the first pass's hand-overs (wide bands, score-0 pairs, thin segments)
        // on their own stream, next to the lane kernel; a second pair of launches after the last pass takes the rest
        BandArgs early = ba;
        early.scratch = b->d_wscr; early.scratch_stride = b->wstride; early.dir_bytes = b->wdir;
        cudaStream_t fb = b->tside[TBAND_INSTANCES];
        cudaEvent_t evFb = b->tev[TBAND_INSTANCES + 1];
        const int wblocks = b->wblocks;
        ssw_batch* bb = b;
        auto after_first_sort = [=](cudaEvent_t sorted) -> cudaError_t {
            cudaError_t e = cudaStreamWaitEvent(fb, sorted, 0);
            if (e != cudaSuccess) return e;
            // (the lists keep growing while these run: each launch works on a snapshot of its count, then the fetch
            // cursor is put back on the snapshot for the closing launches)
            BandArgs x = early;
            int32_t* snap = cnt + 4;
            if ((e = snapshot_count(snap, bb->count2(), fb)) != cudaSuccess) return e;
            x.wl = WorkList{bb->d_idx2, nullptr, snap, bb->cursor2()};
            if ((e = launch_band(1, x, wblocks, fb)) != cudaSuccess) return e;
            if ((e = rewind_cursor(bb->cursor2(), snap, fb)) != cudaSuccess) return e;
            if ((e = snapshot_count(snap + 1, bb->count3(), fb)) != cudaSuccess) return e;
            x.wl = WorkList{bb->d_idx4, nullptr, snap + 1, bb->cursor3()};
            if ((e = launch_band(2, x, wblocks, fb)) != cudaSuccess) return e;
            if ((e = rewind_cursor(bb->cursor3(), snap + 1, fb)) != cudaSuccess) return e;
            return cudaEventRecord(evFb, fb);
        };
        *launches += 2;
        CU_TRY(launch_tband(ta, b->tplan, ls.idx, ls.count, b->n, b->d_tlists + 2 * nn, b->d_tlists + 3 * nn, cnt, cnt + 1, st,
                            b->tside, b->tev, after_first_sort, launches));
        CU_TRY(cudaStreamWaitEvent(st, evFb, 0));
    } else {
        CU_TRY(launch_band(0, ba, b->bblocks, st));
        *launches += 1;
    }
    ba.scratch = b->d_wscr; ba.scratch_stride = b->wstride; ba.dir_bytes = b->wdir;
    ba.wl = WorkList{b->d_idx2, nullptr, b->count2(), b->cursor2()};
    CU_TRY(launch_band(1, ba, b->wblocks, st));
    *launches += 1;
    ba.wl = WorkList{b->d_idx4, nullptr, b->count3(), b->cursor3()};
    CU_TRY(launch_band(2, ba, b->wblocks, st));
    *launches += 1;
    return SSW_OK;
}

extern "C" int ssw_batch_encode_ascii(ssw_batch* b)
{
    if (!b) return SSW_ERR_ARG;
    if (b->ascii_done) return SSW_OK;
    CU_TRY(cudaSetDevice(b->device));
    CU_TRY(encode_ascii(b->d_seqs, b->seq_hi - b->seq_lo, b->stream));
    b->ascii_done = true;
    return SSW_OK;
}

extern "C" int ssw_batch_run(ssw_batch* b)
{
    if (!b) return SSW_ERR_ARG;
    CU_TRY(cudaSetDevice(b->device));
    if (b->n == 0) return SSW_OK;
    cudaStream_t st = b->stream;
    TraceTimer tt("batch_run (enqueue)");
    int launches = 0;
    const BatchView view = b->view();
    const ListSet ls = b->lists();
    CU_TRY(cudaMemsetAsync(b->count2(), 0, 4 * N_LISTS * 4, st));
    const size_t nn3 = ((size_t)std::max(b->n, 1) + 16383) & ~(size_t)16383;       // stride of the two hand-over lists in d_idx3

    // The launches of one stage (one per strip-height list) are independent: they go round robin to side streams,
    // each with scratch of its own, between two events on the batch's stream -- a launch that is down to its last
    // long tasks shares the machine with the next lists instead of holding it (class 1: a few 16 k-column tasks can
    // take milliseconds; thin lists of mixed batches likewise).  fan_pick names stream, scratch and grid of a launch.
    const bool fanning = b->fan0 > 1 || b->fan1 > 1;
    int fan_rr[2] = {0, 0};
    cudaStream_t fs = st;
    unsigned char* fan_scr = nullptr;
    int fan_blocks = 0;
    auto fan_begin = [&]() -> int {
        if (!fanning) return SSW_OK;
        CU_TRY(cudaEventRecord(b->sev[ssw_batch::FAN], st));
        for (auto& x : b->sside) CU_TRY(cudaStreamWaitEvent(x, b->sev[ssw_batch::FAN], 0));
        return SSW_OK;
    };
    auto fan_end = [&]() -> int {
        if (!fanning) return SSW_OK;
        for (int k = 0; k < ssw_batch::FAN; ++k) {
            CU_TRY(cudaEventRecord(b->sev[k], b->sside[k]));
            CU_TRY(cudaStreamWaitEvent(st, b->sev[k], 0));
        }
        return SSW_OK;
    };
    auto fan_pick = [&](int cls, bool serial) {
        fs = st; fan_scr = b->d_sscr[cls]; fan_blocks = b->sblocks[cls];
        if (cls == 0 && b->fan0 > 1) {
            const int k = serial ? 0 : fan_rr[0]++ % b->fan0;
            fs = b->sside[k]; fan_scr = b->d_sscr0x[k];
        } else if (cls == 1 && b->fan1 > 1) {
            const int k = serial ? 0 : fan_rr[1]++ % b->fan1;
            fan_blocks = b->sblocks[1] / 2;
            fs = b->sside[ssw_batch::FAN - 1 - k];                  // (class 0 starts at stream 0, class 1 at the other end)
            fan_scr = b->d_sscr[1] + (size_t)k * fan_blocks * SCORE_WARPS * b->sstride[1];
        } else if (fanning) fs = b->sside[0];
    };

    auto score_args = [&](int cls) {
        ScoreArgs a;
        a.b = view; a.sc = b->sc;
        a.scratch = fan_scr; a.scratch_stride = b->sstride[cls];
        a.off_col = b->off_col[cls]; a.off_bnd = b->off_bnd[cls]; a.off_snap = b->off_snap[cls];
        a.rerun = 0; a.next_idx = nullptr; a.next_base = nullptr; a.next_count = nullptr;
        a.wide_idx = nullptr; a.wide_count = nullptr;
        memset(&a.ck, 0, sizeof a.ck);
        return a;
    };

    // class-1 forward launches: expand the pair list into column-chunk tasks and let the kernel walk those
    int maxScore = 0;
    for (int k = 0; k < 25; ++k) maxScore = std::max<int>(maxScore, b->sc.mat[k]);
    auto chunk_launch = [&](ScoreArgs& a, int kind, int K) -> int {
        if (!b->chunk_cols || !b->long_pairs[kind][K]) return SSW_OK;
        int32_t* cnt = b->d_task_meta + kind * (KMAX + 1) + K;
        int32_t* cur = b->d_task_meta + 2 * (KMAX + 1) + kind * (KMAX + 1) + K;
        CU_TRY(cudaMemsetAsync(cnt, 0, 4, fs));
        CU_TRY(cudaMemsetAsync(cur, 0, 4, fs));
        a.ck.chunk_cols = b->chunk_cols; a.ck.max_match = maxScore;
        a.ck.task_pair = b->d_task + b->task_base[kind][K];
        a.ck.task_c0 = b->d_task + b->task_total + b->task_base[kind][K];
        a.ck.task_c1 = b->d_task + 2 * b->task_total + b->task_base[kind][K];
        a.ck.pair_key = b->d_pair_key; a.ck.pair_left = b->d_pair_left;
        a.ck.col_off = b->d_col_off; a.ck.col_pool = b->d_col_pool;
        CU_TRY(expand_tasks(a.wl, b->long_pairs[kind][K], false, view, b->sc, a.ck, cnt, fs, &launches));
        a.wl = WorkList{nullptr, nullptr, cnt, cur};
        return SSW_OK;
    };

    // ---- forward pass
    CU_TRY(cudaEventRecord(b->ev[0], st));
    CU_TRY(build_lists(0, view, b->sc, LONG_REF_THRESHOLD, ls, st, &launches));
    if (b->n_tiny > 0) {
        TinyArgs ta;
        ta.b = view; ta.sc = b->sc;
        ta.wl = WorkList{ls.idx, ls.base + LIST_TINY, ls.count + LIST_TINY, ls.cursor + LIST_TINY};
        const int blocks = (int)std::min<long long>(((long long)b->n_tiny + 127) / 128, (long long)b->sms * 16);
        CU_TRY(launch_tiny(ta, blocks, st));
        ++launches;
    }
    { const int rc = fan_begin(); if (rc != SSW_OK) return rc; }
    for (int cls = 0; cls < 2; ++cls)
        for (int kind = 0; kind < 2; ++kind)
            for (int K = 1; K <= KMAX; ++K) {
                if (!b->have[cls][kind][K]) continue;
                const int id = list_id(cls, kind, K);
                fan_pick(cls, false);
                ScoreArgs a = score_args(cls);
                a.wl = WorkList{ls.idx, ls.base + id, ls.count + id, ls.cursor + id};
                if (kind == 1) { a.next_idx = b->d_idx2; a.next_base = ls.base + id; a.next_count = b->count2() + id; }
                else { a.next_idx = b->d_idx4; a.next_base = ls.base + id; a.next_count = b->count3() + id; }   // GOTOH overflow -> TRUNC
                a.wide_idx = b->d_idx3 + (size_t)kind * nn3; a.wide_count = b->count2() + LIST_WIDE32 + 2 + kind;
                if (cls == 1) { const int rc = chunk_launch(a, kind, K); if (rc != SSW_OK) return rc; }
                CU_TRY(launch_score(K, kind == 1, false, a, fan_blocks, fs));
                ++launches;
            }
    { const int rc = fan_end(); if (rc != SSW_OK) return rc; }
    CU_TRY(cudaEventRecord(b->ev[1], st));
    { const int rc = fan_begin(); if (rc != SSW_OK) return rc; }
    // ---- pairs whose GOTOH-first pass overflowed 8 bits: the truncated-F pass gives their (word) result.
    // strip heights of the two flavours agree for queries of up to 1024 rows, the only ones guessed GOTOH-first
    for (int cls = 0; cls < 2; ++cls)
        for (int K = 1; K <= KMAX; ++K) {
            if (!b->have_t2[cls][K]) continue;
            const int id = list_id(cls, 0, K);
            fan_pick(cls, false);
            ScoreArgs a = score_args(cls);
            a.wl = WorkList{b->d_idx4, ls.base + id, b->count3() + id, b->cursor3() + id};
            a.rerun = 1;
            a.wide_idx = b->d_idx3 + nn3; a.wide_count = b->count2() + LIST_WIDE32 + 2 + 1;
            if (cls == 1) { const int rc = chunk_launch(a, 0, K); if (rc != SSW_OK) return rc; }      // the GOTOH-first pairs of this K
            CU_TRY(launch_score(K, true, false, a, fan_blocks, fs));
            ++launches;
        }
    // (the two groups below work on different pairs -- a re-run never re-enters the deciding list -- so they share one fan)
    // ---- deciding byte-flavour pass for pairs whose truncated-F pass stayed below the 8-bit limit
    for (int cls = 0; cls < 2; ++cls)
        for (int K = 1; K <= KMAX; ++K) {
            if (!b->have[cls][1][K]) continue;
            const int id = list_id(cls, 1, K);
            fan_pick(cls, false);
            ScoreArgs a = score_args(cls);
            a.wl = WorkList{b->d_idx2, ls.base + id, b->count2() + id, b->cursor2() + id};
            a.rerun = 1;
            if (cls == 1) { const int rc = chunk_launch(a, 1, K); if (rc != SSW_OK) return rc; }      // the TRUNC-first pairs of this K
            CU_TRY(launch_score(K, false, false, a, fan_blocks, fs));
            ++launches;
        }
    { const int rc = fan_end(); if (rc != SSW_OK) return rc; }
    // ---- pairs whose score left the 16-bit comfort zone: 32-bit kernels (rare)
    const int wcls = b->d_sscr[1] ? 1 : 0;
    fs = st; fan_scr = b->d_sscr[wcls];
    for (int kind = 0; kind < 2 && b->d_sscr[wcls]; ++kind) {            // (no scratch = nothing but tiny pairs in the batch)
        ScoreArgs a = score_args(wcls);
        a.wl = WorkList{b->d_idx3 + (size_t)kind * nn3, nullptr, b->count2() + LIST_WIDE32 + 2 + kind, b->cursor2() + LIST_WIDE32 + 2 + kind};
        CU_TRY(launch_score32(kind == 1, false, a, b->sblocks[wcls], st));
        ++launches;
    }
    CU_TRY(cudaEventRecord(b->ev[2], st));
    if (b->sc.flag != 0) {
        // ---- reverse pass: strip height follows the read prefix, so any K up to the forward maximum can occur
        CU_TRY(build_lists(1, view, b->sc, LONG_REF_THRESHOLD, ls, st, &launches));
        if (b->chunk_cols) CU_TRY(cudaMemsetAsync(b->count3(), 0, 2 * N_LISTS * 4, st));
        { const int rc = fan_begin(); if (rc != SSW_OK) return rc; }
        for (int cls = 0; cls < 2; ++cls) {
            if (!b->d_sscr[cls]) continue;
            for (int kind = 0; kind < 2; ++kind) {
                if (kind == 1 && b->sc.go != b->sc.ge) continue;
                for (int K = 1; K <= b->maxK; ++K) {
                    const int id = list_id(cls, kind, K);
                    fan_pick(cls, cls == 1 && b->chunk_cols);          // (long references: the reverse task tables are shared, one list at a time)
                    ScoreArgs a = score_args(cls);
                    a.wl = WorkList{ls.idx, ls.base + id, ls.count + id, ls.cursor + id};
                    if (cls == 1 && b->chunk_cols) {
                        // long references: a bounded first look for the stop column; pairs without one go to a list
                        // (same partition of d_idx4) and are expanded into column-chunk tasks over the whole prefix
                        a.ck.chunk_cols = b->chunk_cols; a.ck.max_match = maxScore;
                        a.next_idx = b->d_idx4; a.next_base = ls.base + id; a.next_count = b->count3() + id;
                        CU_TRY(launch_score(K, kind == 1, true, a, fan_blocks, fs));
                        int32_t* cnt = b->d_task_meta + kind * (KMAX + 1) + K;
                        int32_t* cur = b->d_task_meta + 2 * (KMAX + 1) + kind * (KMAX + 1) + K;
                        CU_TRY(cudaMemsetAsync(cnt, 0, 4, fs));
                        CU_TRY(cudaMemsetAsync(cur, 0, 4, fs));
                        ScoreArgs t = a;
                        t.ck.task_pair = b->d_rtask; t.ck.task_c0 = b->d_rtask + b->rev_task_total;
                        t.ck.task_c1 = b->d_rtask + 2 * b->rev_task_total; t.ck.task_res = b->d_rres;
                        t.ck.pair_key = b->d_pair_key; t.ck.pair_left = b->d_pair_left;
                        const WorkList longList{b->d_idx4, ls.base + id, b->count3() + id, nullptr};
                        CU_TRY(expand_tasks(longList, b->long_total, true, view, b->sc, t.ck, cnt, fs, &launches));
                        t.wl = WorkList{nullptr, nullptr, cnt, cur};
                        CU_TRY(launch_score(K, kind == 1, true, t, fan_blocks, fs));
                        launches += 2;
                        continue;
                    }
                    CU_TRY(launch_score(K, kind == 1, true, a, fan_blocks, fs));
                    ++launches;
                }
            }
        }
        { const int rc = fan_end(); if (rc != SSW_OK) return rc; }
        fs = st; fan_scr = b->d_sscr[wcls];
        for (int kind = 0; kind < 2 && b->d_sscr[wcls]; ++kind) {
            const int id = LIST_WIDE32 + kind;
            ScoreArgs a = score_args(wcls);
            a.wl = WorkList{ls.idx, ls.base + id, ls.count + id, ls.cursor + id};
            CU_TRY(launch_score32(kind == 1, true, a, b->sblocks[wcls], st));
            ++launches;
        }
        CU_TRY(cudaEventRecord(b->ev[3], st));
        // ---- CIGAR pass
        if (!b->no_cigar) { const int rc = enqueue_cigar_stage(b, &launches); if (rc != SSW_OK) return rc; }
    }
    else CU_TRY(cudaEventRecord(b->ev[3], st));
    CU_TRY(cudaEventRecord(b->ev[4], st));
    b->launches = launches;
    return SSW_OK;
}

// Device time of the four stages of the last ssw_batch_run, in ms: forward score pass, deciding
// byte-flavour pass, reverse pass, CIGAR pass (CUDA events on the batch's stream; waits for the run).
extern "C" int ssw_batch_stage_ms(ssw_batch* b, float* ms4)
{
    if (!b || !ms4) return SSW_ERR_ARG;
    CU_TRY(cudaSetDevice(b->device));
    CU_TRY(cudaEventSynchronize(b->ev[4]));
    for (int k = 0; k < 4; ++k) CU_TRY(cudaEventElapsedTime(&ms4[k], b->ev[k], b->ev[k + 1]));
    return SSW_OK;
}

extern "C" int64_t ssw_batch_launch_count(const ssw_batch* b) { return b ? b->launches : 0; }

// Pairs whose direction matrix did not fit the per-warp scratch are re-run one list at a time with a
// scratch sized for the widest band the reference could reach (band < 2*readLen, ssw.c:632).
static int rerun_big_bands(ssw_batch* b, std::vector<int32_t>& big)
{
    cudaStream_t st = b->stream;
    long long need = 0;
    for (int32_t p : big) {
        const PairRec& r = b->h_rec[p];
        // the widest band the doubling loop can reach: it starts at |refLen - readLen| + 1, whatever that is, and goes on
        // while the band is below 2 * readLen (ssw.c:571-632)
        const long long readLen = r.read_end1 - r.read_begin1 + 1, refLen = r.ref_end1 - r.ref_begin1 + 1;
        const long long bw0 = std::llabs(refLen - readLen) + 1;
        const long long bwMax = std::max(bw0, 2 * readLen);
        const long long rowStride = (2 * bwMax + 1 + 15) & ~15LL;
        need = std::max(need, rowStride * readLen + (2 * bwMax + 4) * 8 + 80);
    }
    const long long stride = ((long long)b->bstage * 4 + need + 255) & ~255LL;
    long long warps = std::max<long long>(1, std::min<long long>((6LL << 30) / stride, (long long)big.size()));
    const int blocks = (int)std::max<long long>(1, std::min<long long>((warps + BAND_WARPS - 1) / BAND_WARPS, b->sms));
    unsigned char* scr = nullptr;
    CU_TRY(dev_alloc_t(&scr, (size_t)(stride * blocks * BAND_WARPS), st));
    const ListSet ls = b->lists();
    const int32_t cnt = (int32_t)big.size();
    for (int32_t p : big) b->h_rec[p].status &= ~PS_BAND_SCRATCH;
    std::vector<int32_t> zeros(N_LISTS, 0);
    CU_TRY(cudaMemcpyAsync(ls.idx, big.data(), big.size() * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(ls.cursor, zeros.data(), N_LISTS * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(ls.count, &cnt, 4, cudaMemcpyHostToDevice, st));
    // clear the flag on the device records too
    for (int32_t p : big)
        CU_TRY(cudaMemcpyAsync(&b->d_rec[p].status, &b->h_rec[p].status, 4, cudaMemcpyHostToDevice, st));
    BandArgs ba;
    ba.b = b->view(); ba.sc = b->sc;
    ba.wl = WorkList{ls.idx, nullptr, ls.count, ls.cursor};
    ba.scratch = scr; ba.scratch_stride = stride; ba.dir_bytes = need;
    ba.cigar_stage_cap = b->bstage; ba.cigar_buf = b->d_cigar; ba.cigar_cap = b->cigar_cap; ba.cigar_used = b->d_cigar_used;
    ba.next_idx = b->d_idx2; ba.next_count = b->count2();
    ba.next2_idx = b->d_idx4; ba.next2_count = b->count3();
    CU_TRY(cudaMemsetAsync(b->count2(), 0, 4 * N_LISTS * 4, st));
    CU_TRY(launch_band(0, ba, blocks, st));
    ba.wl = WorkList{b->d_idx2, nullptr, b->count2(), b->cursor2()};
    CU_TRY(launch_band(1, ba, blocks, st));
    ba.wl = WorkList{b->d_idx4, nullptr, b->count3(), b->cursor3()};
    CU_TRY(launch_band(2, ba, blocks, st));
    b->launches += 1;
    b->launches += 2;
    for (int32_t p : big)
        CU_TRY(cudaMemcpyAsync(&b->h_rec[p], &b->d_rec[p], sizeof(PairRec), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    dev_free(scr, st);
    return SSW_OK;
}

// Results of a batch.  `reserve(used, &base)` names the destination of the batch's `used` CIGAR ops and the offset
// that destination has in the caller's buffer (added to every cigar_off); NULL = it does not fit.
template <typename Reserve>
static int fetch_core(ssw_batch* b, ssw_result* out, Reserve reserve, int64_t* cigar_used)
{
    if (!b || (!out && b->n > 0)) return SSW_ERR_ARG;
    CU_TRY(cudaSetDevice(b->device));
    if (cigar_used) *cigar_used = 0;
    if (b->n == 0) return SSW_OK;
    cudaStream_t st = b->stream;
    TraceTimer tt("batch_fetch (total)");
    static_assert(sizeof(PairRec) == sizeof(ssw_result) && offsetof(PairRec, cigar_off) == offsetof(ssw_result, cigar_off) &&
                  offsetof(PairRec, status) == offsetof(ssw_result, status) && offsetof(PairRec, word) == offsetof(ssw_result, word),
                  "device records are copied straight into the caller's result array");
    b->h_rec = reinterpret_cast<PairRec*>(out);
    { TraceTimer t2("  fetch: wait for kernels"); CU_TRY(cudaStreamSynchronize(st)); }
    { TraceTimer t2("  fetch: d2h records");
    CU_TRY(cudaMemcpyAsync(b->h_rec, b->d_rec, (size_t)b->n * sizeof(PairRec), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st)); }
    const bool cigar_stage = b->sc.flag != 0 && !b->no_cigar;
    unsigned long long used = 0;
    int64_t cig_base = 0;
    if (cigar_stage) {
        CU_TRY(cudaMemcpyAsync(&used, b->d_cigar_used, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if ((long long)used > b->cigar_cap && b->cigar_cap < b->cigar_worst) {
            // the estimated CIGAR buffer overflowed: grow it to the worst case and redo the CIGAR pass
            dev_free(b->d_cigar, st);
            b->d_cigar = nullptr;
            b->cigar_cap = b->cigar_worst;
            CU_TRY(dev_alloc_t(&b->d_cigar, (size_t)b->cigar_cap, st));
            CU_TRY(clear_status_bits(b->view(), PS_CIGAR_CAP | PS_BAND_SCRATCH, st));
            int launches = 1;
            { const int rc = enqueue_cigar_stage(b, &launches); if (rc != SSW_OK) return rc; }
            b->launches += launches;
            CU_TRY(cudaMemcpyAsync(b->h_rec, b->d_rec, (size_t)b->n * sizeof(PairRec), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
        // pairs whose direction matrix did not fit the per-warp scratch (after a possible regrow, which repeats the stage)
        std::vector<int32_t> big;
        for (int32_t p = 0; p < b->n; ++p) if (b->h_rec[p].status & PS_BAND_SCRATCH) big.push_back(p);
        if (!big.empty()) { const int rc = rerun_big_bands(b, big); if (rc != SSW_OK) return rc; }
        CU_TRY(cudaMemcpyAsync(&used, b->d_cigar_used, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (cigar_used) *cigar_used = (int64_t)used;
        if ((long long)used > b->cigar_cap) { set_error("internal cigar buffer exhausted"); return SSW_ERR_CIGAR_CAP; }
        if (used > 0) {
            uint32_t* dst = reserve((int64_t)used, &cig_base);
            if (!dst) { set_error("cigar buffer too small"); return SSW_ERR_CIGAR_CAP; }
            TraceTimer t2("  fetch: d2h cigars");
            CU_TRY(cudaMemcpyAsync(dst, b->d_cigar, (size_t)used * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
    }
    // in place: internal stage bits -> public status, CIGAR offsets -> offsets in the caller's buffer
    auto convert = [&](int32_t p0, int32_t p1) {
        for (int32_t p = p0; p < p1; ++p) {
            ssw_result& o = out[p];
            const int32_t st_bits = o.status;
            int32_t pub;
            if (st_bits & (PS_PUNT | PS_UNSUPPORTED | PS_NEED_GOTOH | PS_BAND_SCRATCH | PS_CIGAR_CAP)) pub = SSW_PAIR_UNSUPPORTED;
            else if (st_bits & PS_TRACEBACK_ERR) pub = SSW_PAIR_TRACEBACK_ERR;
            else pub = SSW_PAIR_OK;
            o.status = pub | (st_bits << 8);          // internal stage bits, for diagnostics (see ssw_cuda.h)
            o.cigar_off += cig_base;
        }
    };
    if (b->n < 262144) convert(0, b->n);
    else {
        const int nt = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(convert, (int32_t)((long long)b->n * t / nt), (int32_t)((long long)b->n * (t + 1) / nt));
        for (auto& x : th) x.join();
    }
    return SSW_OK;
}

extern "C" int ssw_batch_fetch(ssw_batch* b, ssw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used)
{
    return fetch_core(b, out, [&](int64_t used, int64_t* base) -> uint32_t* {
        *base = 0;
        return (cigar_buf && used <= cigar_cap) ? cigar_buf : nullptr;
    }, cigar_used);
}

static int error_code_of_create()
{
    return g_last_error.find("not supported") != std::string::npos ? SSW_ERR_UNSUPPORTED
         : (g_last_error.find("no usable CUDA") != std::string::npos ? SSW_ERR_NODEVICE
         : (g_last_error.find("cuda") != std::string::npos ? SSW_ERR_CUDA : SSW_ERR_ARG));
}

// One-shot call on host buffers, on one or several devices of the box (pairs are
// independent, so the batch shards with no exchange step and no collective).
//
// The pair list is cut into contiguous chunks of SSW_CUDA_CHUNK pairs (default 262144, fewer when that would
// leave a device with less than ~8 chunks); one host thread per device pulls chunks from a shared counter --
// longest-queue-first balancing without a cost model -- and runs each as its own device batch on alternating
// streams: the host-to-device copy of chunk k+1 and the device-to-host copy of chunk k-1 overlap the kernels of
// chunk k (with pinned caller memory; pageable memory still works, the copies then serialise).  A chunk uploads
// only the byte range of `seqs` its pairs reference, so a pair-major layout moves every byte once and to one
// device only.  Results land at the pairs' own indices (the host-side gather is the addressing); CIGAR space
// in the caller's buffer is reserved with one atomic add per chunk, so cigar_off values are absolute.
struct MultiJob {
    int32_t n_pairs, chunk;
    const int8_t* seqs; int64_t seqs_len;
    const int64_t* q_off; const int32_t* q_len; const int64_t* r_off; const int32_t* r_len; const int32_t* mask_len;
    const ssw_scoring* scoring;
    bool packed = false;
    ssw_result* out; uint32_t* cigar_buf; int64_t cigar_cap;
    std::atomic<int64_t> next_chunk{0};
    std::atomic<int64_t> cig_cursor{0};
    std::atomic<int> rc{SSW_OK};
    std::mutex err_mu; std::string err;
    void fail(int code, const std::string& msg) { int ok = SSW_OK; if (rc.compare_exchange_strong(ok, code)) { std::lock_guard<std::mutex> g(err_mu); err = msg; } }
};

static void multi_worker(MultiJob* J, int device, int slots)
{
    constexpr int MAX_SLOTS = 4;
    ssw_batch* slot[MAX_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    int32_t slot_p0[MAX_SLOTS] = {0, 0, 0, 0};
    auto finish = [&](int sidx) {
        ssw_batch* b = slot[sidx];
        if (!b) return;
        if (J->rc.load() == SSW_OK) {
            int64_t used = 0;
            const int r = fetch_core(b, J->out + slot_p0[sidx], [&](int64_t n, int64_t* base) -> uint32_t* {
                const int64_t at = J->cig_cursor.fetch_add(n);
                *base = at;
                return (J->cigar_buf && at + n <= J->cigar_cap) ? J->cigar_buf + at : nullptr;
            }, &used);
            if (r != SSW_OK) J->fail(r, g_last_error);
        }
        ssw_batch_destroy(b);
        slot[sidx] = nullptr;
    };
    const int64_t n_chunks = ((int64_t)J->n_pairs + J->chunk - 1) / J->chunk;
    int k = 0;
    for (;; ++k) {
        if (J->rc.load() != SSW_OK) break;
        const int64_t c = J->next_chunk.fetch_add(1);
        if (c >= n_chunks) break;
        const int32_t p0 = (int32_t)(c * J->chunk);
        const int32_t cnt = std::min<int32_t>(J->chunk, J->n_pairs - p0);
        const int sidx = k % slots;
        finish(sidx);                                   // (normally already drained below)
        ssw_batch* b = batch_create_impl(device, nullptr, cnt, J->seqs, J->seqs_len, J->q_off + p0, J->q_len + p0, J->r_off + p0,
                                         J->r_len + p0, J->mask_len ? J->mask_len + p0 : nullptr, J->scoring, J->packed);
        if (!b) { J->fail(error_code_of_create(), g_last_error); break; }
        slot[sidx] = b; slot_p0[sidx] = p0;
        const int r = ssw_batch_run(b);
        if (r != SSW_OK) { J->fail(r, g_last_error); break; }
        finish((k + 1) % slots);                        // the oldest chunk in flight on this device
    }
    for (int j = 1; j <= slots; ++j) finish((k + j) % slots);      // drain, oldest first
    for (int sidx = 0; sidx < MAX_SLOTS; ++sidx) finish(sidx);
}

static int align_multi_impl(const int* devices, int n_devices, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                            const int64_t* q_off, const int32_t* q_len, const int64_t* r_off, const int32_t* r_len,
                            const int32_t* mask_len, const ssw_scoring* scoring, ssw_result* out, uint32_t* cigar_buf,
                            int64_t cigar_cap, int64_t* cigar_used, bool packed);

extern "C" int ssw_align_batch_multi(const int* devices, int n_devices, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                                     const int64_t* q_off, const int32_t* q_len, const int64_t* r_off, const int32_t* r_len,
                                     const int32_t* mask_len, const ssw_scoring* scoring, ssw_result* out, uint32_t* cigar_buf,
                                     int64_t cigar_cap, int64_t* cigar_used)
{
    return align_multi_impl(devices, n_devices, n_pairs, seqs, seqs_len, q_off, q_len, r_off, r_len, mask_len, scoring, out, cigar_buf,
                            cigar_cap, cigar_used, false);
}

extern "C" int ssw_align_batch_multi_packed(const int* devices, int n_devices, int32_t n_pairs, const uint8_t* packed, int64_t n_bases,
                                            const int64_t* q_off, const int32_t* q_len, const int64_t* r_off, const int32_t* r_len,
                                            const int32_t* mask_len, const ssw_scoring* scoring, ssw_result* out, uint32_t* cigar_buf,
                                            int64_t cigar_cap, int64_t* cigar_used)
{
    return align_multi_impl(devices, n_devices, n_pairs, reinterpret_cast<const int8_t*>(packed), n_bases, q_off, q_len, r_off, r_len,
                            mask_len, scoring, out, cigar_buf, cigar_cap, cigar_used, true);
}

static int align_multi_impl(const int* devices, int n_devices, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                            const int64_t* q_off, const int32_t* q_len, const int64_t* r_off, const int32_t* r_len,
                            const int32_t* mask_len, const ssw_scoring* scoring, ssw_result* out, uint32_t* cigar_buf,
                            int64_t cigar_cap, int64_t* cigar_used, bool packed)
{
    if (cigar_used) *cigar_used = 0;
    if (n_pairs <= 0) return n_pairs == 0 ? SSW_OK : SSW_ERR_ARG;
    if (!devices || n_devices <= 0 || n_devices > 64) { set_error("ssw_align_batch_multi: invalid device list"); return SSW_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no usable CUDA device (libssw_cuda has no CPU path)"); return SSW_ERR_NODEVICE; }
    for (int d = 0; d < n_devices; ++d)
        if (devices[d] < 0 || devices[d] >= ndev) { set_error("ssw_align_batch_multi: device " + std::to_string(devices[d]) + " does not exist"); return SSW_ERR_NODEVICE; }
    int32_t chunk = 262144;
    if (const char* e = getenv("SSW_CUDA_CHUNK")) { const long v = atol(e); if (v > 0 && v < (1L << 30)) chunk = (int32_t)v; }
    else {
        // about eight chunks per device: enough for the shared queue to balance and for the copies of one chunk to hide
        // behind the kernels of another; never so small that a chunk stops filling a GPU (mixed batches spread over
        // many kernel instances), never beyond a million pairs
        const int64_t want = (int64_t)n_pairs / ((int64_t)n_devices * 8);
        chunk = (int32_t)std::max<int64_t>(n_devices > 1 ? 65536 : 262144, std::min<int64_t>(1 << 20, want));
    }
    int slots = 4;
    if (const char* e = getenv("SSW_CUDA_SLOTS")) { const long v = atol(e); if (v >= 2 && v <= 4) slots = (int)v; }
    MultiJob J;
    J.n_pairs = n_pairs; J.chunk = chunk; J.seqs = seqs; J.seqs_len = seqs_len; J.q_off = q_off; J.q_len = q_len;
    J.r_off = r_off; J.r_len = r_len; J.mask_len = mask_len; J.scoring = scoring; J.out = out; J.cigar_buf = cigar_buf; J.cigar_cap = cigar_cap;
    J.packed = packed;
    if (n_devices == 1) multi_worker(&J, devices[0], slots);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < n_devices; ++d) th.emplace_back(multi_worker, &J, devices[d], slots);
        for (auto& t : th) t.join();
    }
    if (cigar_used) *cigar_used = J.cig_cursor.load();
    if (J.rc.load() != SSW_OK) { set_error(J.err); return J.rc.load(); }
    return SSW_OK;
}

extern "C" int ssw_align_batch(int device, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len, const int64_t* q_off,
                               const int32_t* q_len, const int64_t* r_off, const int32_t* r_len, const int32_t* mask_len,
                               const ssw_scoring* scoring, ssw_result* out, uint32_t* cigar_buf, int64_t cigar_cap,
                               int64_t* cigar_used)
{
    return ssw_align_batch_multi(&device, 1, n_pairs, seqs, seqs_len, q_off, q_len, r_off, r_len, mask_len, scoring, out,
                                 cigar_buf, cigar_cap, cigar_used);
}

extern "C" void ssw_encode_dna(const char* ascii, int64_t len, int8_t* codes)
{
    static int8_t lut[256];
    static bool init = false;
    if (!init) {
        memset(lut, 4, sizeof lut);
        const char* bases = "ACGTN";
        for (int k = 0; k < 5; ++k) { lut[(unsigned char)bases[k]] = (int8_t)k; lut[(unsigned char)(bases[k] + 32)] = (int8_t)k; }
        init = true;
    }
    for (int64_t k = 0; k < len; ++k) codes[k] = lut[(unsigned char)ascii[k]];
}

// ---- batched edit distance (CIRI_long/utils.py:153-159) -------------------
// One call: upload the referenced bytes, classify the pairs by the length of their shorter string into
// the kernel instances of edit_distance.cu, run them on one stream, copy the distances back.
extern "C" int ssw_cuda_edit_distance_batch(int device, int32_t n_pairs, const uint8_t* seqs, int64_t seqs_len,
                                            const int64_t* x_off, const int32_t* x_len,
                                            const int64_t* y_off, const int32_t* y_len, int32_t* out)
{
    if (n_pairs < 0 || (n_pairs > 0 && (!seqs || !x_off || !x_len || !y_off || !y_len || !out))) {
        set_error("ssw_cuda_edit_distance_batch: null argument");
        return SSW_ERR_ARG;
    }
    if (n_pairs == 0) return SSW_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("ssw_cuda_edit_distance_batch: no CUDA device (this library has no CPU path)");
        return SSW_ERR_NODEVICE;
    }
    CU_TRY(cudaSetDevice(device));
    // byte range referenced by the batch, symbols that occur in it, work lists per kernel instance
    long long lo = seqs_len, hi = 0;
    std::vector<int32_t> lists[6];
    long long maxText5 = 0;
    bool seen[256] = {false};
    for (int32_t p = 0; p < n_pairs; ++p) {
        const long long xl = x_len[p], yl = y_len[p];
        if (xl < 0 || yl < 0 || x_off[p] < 0 || y_off[p] < 0 || x_off[p] + xl > seqs_len || y_off[p] + yl > seqs_len) {
            set_error("ssw_cuda_edit_distance_batch: pair " + std::to_string(p) + " lies outside the sequence buffer");
            return SSW_ERR_ARG;
        }
        if (xl) { lo = std::min<long long>(lo, x_off[p]); hi = std::max<long long>(hi, x_off[p] + xl); }
        if (yl) { lo = std::min<long long>(lo, y_off[p]); hi = std::max<long long>(hi, y_off[p] + yl); }
        const long long mm = std::min(xl, yl);
        const int kind = mm <= 32 ? 0 : mm <= 64 ? 1 : mm <= 128 ? 2 : mm <= 256 ? 3 : mm <= 512 ? 4 : 5;
        lists[kind].push_back(p);
        if (kind == 5) maxText5 = std::max(maxText5, std::max(xl, yl));
    }
    if (hi < lo) { lo = 0; hi = 0; }
    for (long long k = lo; k < hi; ++k) seen[seqs[k]] = true;
    EditArgs ea;
    int nsym = 0;
    for (int b = 0; b < 256; ++b) {
        ea.code[b] = 0;
        if (seen[b]) {
            if (nsym == ED_MAXSYM) {
                set_error("ssw_cuda_edit_distance_batch: more than 16 distinct symbols in the batch");
                return SSW_ERR_UNSUPPORTED;
            }
            ea.code[b] = (unsigned char)nsym++;
        }
    }
    cudaStream_t st = nullptr;
    CU_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int sms = 0;
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    unsigned char* d_seqs = nullptr;
    int64_t *d_xoff = nullptr, *d_yoff = nullptr;
    int32_t *d_xlen = nullptr, *d_ylen = nullptr, *d_out = nullptr, *d_idx = nullptr;
    signed char* d_carry = nullptr;
    int rc = SSW_OK;
    auto body = [&]() -> int {
        CU_TRY(dev_alloc_t(&d_seqs, (size_t)(hi - lo), st));
        CU_TRY(dev_alloc_t(&d_xoff, n_pairs, st)); CU_TRY(dev_alloc_t(&d_yoff, n_pairs, st));
        CU_TRY(dev_alloc_t(&d_xlen, n_pairs, st)); CU_TRY(dev_alloc_t(&d_ylen, n_pairs, st));
        CU_TRY(dev_alloc_t(&d_out, n_pairs, st)); CU_TRY(dev_alloc_t(&d_idx, n_pairs, st));
        if (hi > lo) CU_TRY(cudaMemcpyAsync(d_seqs, seqs + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(d_xoff, x_off, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(d_yoff, y_off, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(d_xlen, x_len, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(d_ylen, y_len, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, st));
        ea.seqs = d_seqs - lo; ea.x_off = d_xoff; ea.x_len = d_xlen; ea.y_off = d_yoff; ea.y_len = d_ylen;
        ea.out = d_out; ea.carry = nullptr; ea.carry_stride = 0;
        long long at = 0;
        for (int kind = 0; kind < 6; ++kind) {
            const long long cnt = (long long)lists[kind].size();
            if (!cnt) continue;
            CU_TRY(cudaMemcpyAsync(d_idx + at, lists[kind].data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, st));
            ea.idx = d_idx + at; ea.count = (int32_t)cnt;
            long long blocks;
            if (kind < 2) blocks = std::min<long long>((cnt + EDIT_THREADS - 1) / EDIT_THREADS, (long long)sms * 16);
            else {
                const int groupsPerBlock = (EDIT_THREADS / 32) * (32 / (4 << (kind - 2)));
                blocks = std::min<long long>((cnt + groupsPerBlock - 1) / groupsPerBlock, (long long)sms * 8);
            }
            blocks = std::max<long long>(blocks, 1);
            if (kind == 5) {
                ea.carry_stride = (maxText5 + 255) & ~255LL;
                CU_TRY(dev_alloc_t(&d_carry, (size_t)(blocks * (EDIT_THREADS / 32) * ea.carry_stride), st));
                ea.carry = d_carry;
            }
            CU_TRY(launch_edit(kind, ea, (int)blocks, st));
            at += cnt;
        }
        CU_TRY(cudaMemcpyAsync(out, d_out, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        return SSW_OK;
    };
    rc = body();
    dev_free(d_seqs, st); dev_free(d_xoff, st); dev_free(d_yoff, st); dev_free(d_xlen, st); dev_free(d_ylen, st);
    dev_free(d_out, st); dev_free(d_idx, st); dev_free(d_carry, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int ssw_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int ssw_cuda_dpx_peak(int device, double* lane_instr_per_s, double* sm_clock_mhz)
{
    CU_TRY(cudaSetDevice(device));
    int khz = 0;
    CU_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
    CU_TRY(dpx_peak_probe(lane_instr_per_s, 0));
    return SSW_OK;
}

// ---------------------------------------------------------------------------------------------------
// legacy six-symbol ABI (ssw.h): a batch of one

struct _profile {
    const int8_t* read;
    const int8_t* mat;
    int32_t readLen;
    int32_t n;
    int8_t score_size;
};

extern "C" s_profile* ssw_init(const int8_t* read, const int32_t readLen, const int8_t* mat, const int32_t n,
                               const int8_t score_size)
{
    s_profile* p = (s_profile*)calloc(1, sizeof(struct _profile));
    if (!p) return nullptr;
    p->read = read; p->mat = mat; p->readLen = readLen; p->n = n; p->score_size = score_size;
    return p;
}

extern "C" void init_destroy(s_profile* p) { free(p); }

extern "C" s_align* ssw_align(const s_profile* prof, const int8_t* ref, int32_t refLen, const uint8_t weight_gapO,
                              const uint8_t weight_gapE, const uint8_t flag, const uint16_t filters,
                              const int32_t filterd, const int32_t maskLen)
{
    if (!prof || !prof->read || !prof->mat) {
        fprintf(stderr, "Please call the function ssw_init before ssw_align.\n");
        return nullptr;
    }
    if (maskLen < 15)
        fprintf(stderr, "When maskLen < 15, the function ssw_align doesn't return 2nd best alignment information.\n");
    if (prof->n != 5 || prof->score_size != 2) {
        fprintf(stderr, "libssw_cuda: only the 5-letter DNA alphabet with score_size 2 (the ssw_wrap.py call) is implemented on the device.\n");
        return nullptr;
    }
    if (prof->readLen <= 0 || refLen <= 0) {
        fprintf(stderr, "libssw_cuda: empty read or reference.\n");
        return nullptr;
    }
    ssw_scoring sc;
    memset(&sc, 0, sizeof sc);
    memcpy(sc.mat, prof->mat, 25);
    sc.gap_open = weight_gapO; sc.gap_extend = weight_gapE; sc.flag = flag; sc.filters = filters; sc.filterd = filterd;
    std::vector<int8_t> seqs((size_t)prof->readLen + refLen);
    memcpy(seqs.data(), prof->read, prof->readLen);
    memcpy(seqs.data() + prof->readLen, ref, refLen);
    const int64_t q_off = 0, r_off = prof->readLen;
    const int32_t q_len = prof->readLen, r_len = refLen, mask = maskLen;
    ssw_result res;
    std::vector<uint32_t> cig(2 * (size_t)prof->readLen + 16);
    int64_t used = 0;
    int legacy_dev = 0;                                   // SSW_CUDA_DEVICE: which GPU serves the per-call ABI in this process
    if (const char* e = getenv("SSW_CUDA_DEVICE")) legacy_dev = atoi(e);
    const int rc = ssw_align_batch(legacy_dev, 1, seqs.data(), (int64_t)seqs.size(), &q_off, &q_len, &r_off, &r_len, &mask, &sc,
                                   &res, cig.data(), (int64_t)cig.size(), &used);
    if (rc != SSW_OK) {
        fprintf(stderr, "libssw_cuda: %s\n", g_last_error.c_str());
        return nullptr;
    }
    // The traceback left the band (ssw.c:642-673 then indexes direction bytes no band pass wrote and the
    // reference's CIGAR is whatever that heap memory yields): score and coordinates are exact and are returned,
    // the CIGAR is empty (cigar == NULL, cigarLen == 0) instead of undefined.
    if ((res.status & 0xff) == SSW_PAIR_TRACEBACK_ERR) res.cigar_len = 0;
    else if ((res.status & 0xff) != SSW_PAIR_OK) {
        fprintf(stderr, "libssw_cuda: this pair needs a code path the device library does not provide.\n");
        return nullptr;
    }
    s_align* r = (s_align*)calloc(1, sizeof(s_align));
    r->score1 = (uint16_t)res.score1; r->score2 = (uint16_t)res.score2;
    r->ref_begin1 = res.ref_begin1; r->ref_end1 = res.ref_end1;
    r->read_begin1 = res.read_begin1; r->read_end1 = res.read_end1; r->ref_end2 = res.ref_end2;
    r->cigar = nullptr; r->cigarLen = 0;
    if (res.cigar_len > 0) {
        r->cigar = (uint32_t*)malloc((size_t)res.cigar_len * 4);
        memcpy(r->cigar, cig.data() + res.cigar_off, (size_t)res.cigar_len * 4);
        r->cigarLen = res.cigar_len;
    }
    return r;
}

extern "C" void align_destroy(s_align* a)
{
    if (!a) return;
    free(a->cigar);
    free(a);
}

extern "C" char cigar_int_to_op(uint32_t cigar_int)
{
    static const char ops[] = "MIDNSHP=X";
    const uint32_t code = cigar_int & 0xfU;
    return code < 9 ? ops[code] : 'M';
}

extern "C" uint32_t cigar_int_to_len(uint32_t cigar_int) { return cigar_int >> 4; }
