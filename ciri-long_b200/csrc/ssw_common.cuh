// ssw_common.cuh -- device-side records and launch plumbing shared by the kernels of libssw_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sswb {

constexpr int WARP = 32;
constexpr int SCORE_WARPS = 24;                 // most warps a CTA of the score-pass kernels can have (scratch is sized for it)
// warps per CTA by strip height: fewer rows per strip need fewer registers, and the DPX dependency chains
// want as many warps per scheduler as the register file allows (one CTA per SM: the LUT takes 80 KB)
__host__ __device__ constexpr int score_warps(int K) { return K < 0 ? 24 : 16; }
constexpr int SCORE32_WARPS = 8;                // warps per CTA of the 32-bit score kernels
constexpr int VSTRIPS = 64;                     // virtual strips per warp: 32 lanes x 2 packed halves
constexpr int KMAX = 16;                        // max query rows per strip -> 1024 rows per tile
constexpr int LUT_ENTRIES = 625;                // (ref pair 25) x (query pair 25)
constexpr int LUT_BYTES = LUT_ENTRIES * 32 * 4; // lane-replicated: bank == lane, conflict free
constexpr int RP_STRIDE = 25 * 128;             // bytes between consecutive ref-pair rows of the LUT
constexpr int RP_CHUNK = 512;                   // wavefront steps per staged chunk of ref-pair codes
constexpr int RP_WINDOW = RP_CHUNK + 64;        // bytes of shared memory per warp for that chunk (+31 lanes of skew, +1 look-ahead)
constexpr int SCORE_SMEM_BYTES = LUT_BYTES + SCORE_WARPS * RP_WINDOW;
constexpr unsigned S16X2_MIN = 0x80008000u;
constexpr int TRUNC_GATE = -16384;              // "minus infinity" that cannot wrap s16 when added to a score
constexpr int TRUNC_SCORE_LIMIT = 16000;        // pairs scoring above this leave the truncated-F fast path
constexpr int S16_SCORE_LIMIT = 32767 - 128;    // pairs scoring above this leave the s16 fast path (ssw.c:442 saturates)

// per-pair internal status bits (device)
enum : int32_t {
    PS_NEED_GOTOH = 1,      // truncated-F pass scored below the 8-bit limit: the byte flavour decides (ssw.c:805-809)
    PS_PUNT = 2,            // needs the exact lane-model kernel (saturation, reverse pass exceeded score1, ...)
    PS_TRACEBACK_ERR = 4,   // traceback left the band / undefined direction (reference returns NULL)
    PS_BAND_SCRATCH = 8,    // direction matrix did not fit the per-warp scratch
    PS_CIGAR_CAP = 16,      // cigar output buffer exhausted
    PS_UNSUPPORTED = 32,
    PS_WIDE32 = 64,         // scores near the 16-bit range: handled by the 32-bit score kernels
    PS_REV_DONE = 128       // begin coordinates already computed (tiny pairs: ssw_tiny.cu does both passes)
};

// one record per pair, device resident (mirrors ssw_result + internals)
struct PairRec {
    int32_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    int32_t cigar_len;
    long long cigar_off;
    int32_t status;
    int32_t word;
};

struct BatchView {
    const int8_t* seqs;
    const long long* q_off;
    const int32_t* q_len;
    const long long* r_off;
    const int32_t* r_len;
    const int32_t* mask_len;
    PairRec* rec;
    int32_t n_pairs;
};

struct Scoring {
    int8_t mat[25];
    int32_t go, ge, bias;
    int32_t flag, filters, filterd;
};

// A device work list: indices into the batch + a device-side count + an atomic cursor.
struct WorkList {
    const int32_t* idx;     // index array shared by all lists of a stage
    const int32_t* base;    // device pointer: first entry of this list inside idx (NULL = 0)
    const int32_t* count;   // device pointer: number of valid entries
    int32_t* cursor;        // device pointer: atomic fetch counter (zeroed before launch)
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// DPX wrappers (sm_90+: single VIADDMNMX / VIMNMX / VIMNMX3 instructions)
__device__ __forceinline__ unsigned addmax(unsigned a, unsigned b, unsigned c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ unsigned addmax_relu(unsigned a, unsigned b, unsigned c) { return __viaddmax_s16x2_relu(a, b, c); }
__device__ __forceinline__ unsigned max_relu(unsigned a, unsigned b) { return __vimax_s16x2_relu(a, b); }
__device__ __forceinline__ unsigned max3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }

// shared-memory loads through 32-bit window addresses, and an address multiply-add that stays on the FMA
// pipe (the DPX instructions own the ALU pipe in the score kernels)
__device__ __forceinline__ unsigned lds_u32(unsigned addr) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ unsigned lds_u8(unsigned addr) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ unsigned mad_u32(unsigned a, unsigned b, unsigned c) { unsigned v; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(a), "r"(b), "r"(c)); return v; }

// prmt.b32 with the full selector (bit 3 of a nibble replicates the sign of the selected byte); the CUDA
// intrinsic __byte_perm masks the selector with 0x7777 and loses that mode
__device__ __forceinline__ unsigned prmt_raw(unsigned a, unsigned b, unsigned sel) { unsigned v; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(a), "r"(b), "r"(sel)); return v; }

__device__ __forceinline__ unsigned pack2(int lo, int hi) { return (unsigned)(lo & 0xffff) | ((unsigned)hi << 16); }
__device__ __forceinline__ int lo16(unsigned v) { return (int)(short)(v & 0xffff); }
__device__ __forceinline__ int hi16(unsigned v) { return (int)(short)(v >> 16); }

}  // namespace sswb
