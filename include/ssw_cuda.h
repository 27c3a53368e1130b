/*
 * ssw_cuda.h -- C ABI of libssw_cuda.so: the B200-native replacement for CIRI-long's libssw.so.
 *
 * Two groups of entry points:
 *
 *  (1) The six legacy symbols of the reference library, with identical signatures, struct layout and
 *      ownership rules, so that the reference's ctypes binding (libs/striped_smith_waterman/
 *      ssw_wrap.py:54-72, 278-288) works against this library unchanged.  Each call is a device batch
 *      of one pair.
 *
 *  (2) The batched interface the CIRI-long call sites are moved to (find_bsj.py:182-233,
 *      collapse.py:156-265, 372-387): many independent (query, reference) pairs in one call, as
 *      struct-of-arrays over a concatenated int8 code buffer (A C G T N -> 0..4, ssw_wrap.py:50).
 *
 * All pointers are plain host pointers unless stated otherwise; no torch / CUDA types appear.  Nothing
 * here aborts the process: errors are reported through return codes (and one line on stderr for the
 * legacy calls, like the reference, ssw.c:810-821).  There is no CPU implementation behind these
 * entry points: if no CUDA device is usable they fail.
 */
#ifndef SSW_CUDA_H
#define SSW_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * (1) legacy ABI -- replaces libs/striped_smith_waterman/ssw.h
 * ---------------------------------------------------------------------------------------------- */

/* ssw.h:24-26 (opaque).  Borrows `read` and `mat` until init_destroy, exactly like ssw.c:766-767. */
struct _profile;
typedef struct _profile s_profile;

/* ssw.h:42-52.  sizeof == 40 on LP64, offsets 0,2,4,8,12,16,20,24,32 (mirrored by CAlignRes,
 * ssw_wrap.py:29-37).  `cigar` is malloc'ed by the library and released by align_destroy. */
typedef struct {
    uint16_t score1;
    uint16_t score2;
    int32_t ref_begin1;
    int32_t ref_end1;
    int32_t read_begin1;
    int32_t read_end1;
    int32_t ref_end2;
    uint32_t* cigar;
    int32_t cigarLen;
} s_align;

/* ssw.h:72 / ssw.c:750-771 */
s_profile* ssw_init(const int8_t* read, const int32_t readLen, const int8_t* mat, const int32_t n,
                    const int8_t score_size);
/* ssw.h:77 / ssw.c:773-777 */
void init_destroy(s_profile* p);
/* ssw.h:112-120 / ssw.c:779-869.  Returns NULL on error (one line on stderr). */
s_align* ssw_align(const s_profile* prof, const int8_t* ref, int32_t refLen, const uint8_t weight_gapO,
                   const uint8_t weight_gapE, const uint8_t flag, const uint16_t filters,
                   const int32_t filterd, const int32_t maskLen);
/* ssw.h:125 / ssw.c:871-874 */
void align_destroy(s_align* a);
/* ssw.h:176 / ssw.c:876-896 */
char cigar_int_to_op(uint32_t cigar_int);
/* ssw.h:182 / ssw.c:898-902 */
uint32_t cigar_int_to_len(uint32_t cigar_int);

/* ------------------------------------------------------------------------------------------------
 * (2) batched ABI (new)
 * ---------------------------------------------------------------------------------------------- */

enum {
    SSW_OK = 0,
    SSW_ERR_CUDA = -1,          /* CUDA runtime error (message via ssw_cuda_last_error) */
    SSW_ERR_ARG = -2,           /* invalid argument */
    SSW_ERR_CIGAR_CAP = -3,     /* cigar buffer too small; *cigar_used holds the required size */
    SSW_ERR_UNSUPPORTED = -4,   /* scoring scheme outside what the device kernels implement */
    SSW_ERR_NODEVICE = -5
};

/* per-pair status: low byte of ssw_result.status (bits 8.. carry internal stage flags for diagnostics) */
enum {
    SSW_PAIR_OK = 0,
    SSW_PAIR_TRACEBACK_ERR = 1, /* the traceback left the band: the reference reads direction bytes it never wrote
                                   (ssw.c:642-673) and returns an undefined CIGAR or NULL ("Trace back error",
                                   ssw.c:674-682).  Score and coordinates are exact; cigar_len is 0. */
    SSW_PAIR_UNSUPPORTED = 2    /* pair needs a path this build does not provide (reported, never guessed) */
};

/* One result per pair: the seven s_align fields (same meaning, ssw.h:28-41) with the CIGAR pointer
 * replaced by an (offset, length) window into the caller's cigar buffer. */
typedef struct {
    int32_t score1;
    int32_t score2;
    int32_t ref_begin1;
    int32_t ref_end1;
    int32_t read_begin1;
    int32_t read_end1;
    int32_t ref_end2;
    int32_t cigar_len;
    int64_t cigar_off;
    int32_t status;
    int32_t word;               /* 1 if the 16-bit flavour produced the result (ssw.c:806-809) */
} ssw_result;

/* Scoring: mat is the n*n substitution matrix of ssw_init (n must be 5 for the batched path: the
 * A C G T N alphabet of ssw_wrap.py:146-159); gaps are absolute values as in ssw_align. */
typedef struct {
    int8_t mat[25];
    uint8_t gap_open;
    uint8_t gap_extend;
    uint8_t flag;               /* ssw_align flag: 0 = score + end only, 1 = begin + CIGAR too */
    uint8_t _pad;
    uint16_t filters;
    uint16_t _pad2;
    int32_t filterd;
} ssw_scoring;

typedef struct ssw_batch ssw_batch;   /* opaque: a batch resident in device memory */

/* Upload a batch to `device`.  mask_len may be NULL (the wrapper's rule, ssw_wrap.py:196-199:
 * len(query)//2 if len(query) > 30 else 15).  `stream` is a cudaStream_t passed as void* (NULL: the
 * library creates its own non-blocking stream).  Returns NULL on error. */
ssw_batch* ssw_batch_create(int device, void* stream, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                            const int64_t* q_off, const int32_t* q_len,
                            const int64_t* r_off, const int32_t* r_len,
                            const int32_t* mask_len, const ssw_scoring* scoring);
/* Enqueue all kernels of the hot path on the batch's stream (asynchronous; inputs stay resident,
 * results stay on the device).  May be called repeatedly on the same batch. */
int ssw_batch_run(ssw_batch* b);
/* The batch was created from ASCII letters instead of codes: convert them on the device (A C G T N in either
 * case -> 0..4, anything else 4; the encode of ssw_wrap.py:234-252).  Call once, before ssw_batch_run. */
int ssw_batch_encode_ascii(ssw_batch* b);
/* Wait for the stream, copy results and CIGARs to host.  cigar_buf may be NULL when flag == 0. */
int ssw_batch_fetch(ssw_batch* b, ssw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used);
/* Number of kernel launches the last ssw_batch_run enqueued. */
int64_t ssw_batch_launch_count(const ssw_batch* b);
/* Device time (ms, CUDA events on the batch's stream) of the four stages of the last ssw_batch_run:
 * forward score pass, deciding byte-flavour pass, reverse pass, CIGAR pass.  Waits for the run. */
int ssw_batch_stage_ms(ssw_batch* b, float* ms4);
void ssw_batch_destroy(ssw_batch* b);

/* One-shot convenience: create + run + fetch + destroy on one device. */
int ssw_align_batch(int device, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                    const int64_t* q_off, const int32_t* q_len,
                    const int64_t* r_off, const int32_t* r_len,
                    const int32_t* mask_len, const ssw_scoring* scoring,
                    ssw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used);

/* The same call over several devices of one box (SURVEY.md section 7's `devices, n_devices` signature; section 8e:
 * pairs are independent, no collective).  The pair list is cut into contiguous chunks that one host thread per
 * device pulls from a shared queue; every chunk uploads only the bytes its pairs reference, results land at the
 * pairs' own indices, cigar_off values are absolute offsets into cigar_buf.  Results are bit-identical to the
 * single-device call.  ssw_align_batch(device, ...) is this call with one device. */
int ssw_align_batch_multi(const int* devices, int n_devices, int32_t n_pairs, const int8_t* seqs, int64_t seqs_len,
                          const int64_t* q_off, const int32_t* q_len,
                          const int64_t* r_off, const int32_t* r_len,
                          const int32_t* mask_len, const ssw_scoring* scoring,
                          ssw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used);

/* 4-bit packed input (north_star: packed bases, 128-bit loads; SURVEY.md section 0 item 6: N needs its own code, so
 * 4 bits, not 2).  `packed` holds two bases per byte, low nibble first, codes 0..4 (A C G T N; anything above 4
 * counts as N); offsets and lengths count BASES into that buffer.  Half the bytes cross PCIe; the device expands
 * them with 128-bit loads/stores into the one-code-per-byte layout the kernels read.  Results are identical to the
 * unpacked calls.  ssw_pack_dna4 is the host-side packer (codes -> nibbles). */
ssw_batch* ssw_batch_create_packed(int device, void* stream, int32_t n_pairs, const uint8_t* packed, int64_t n_bases,
                                   const int64_t* q_off, const int32_t* q_len,
                                   const int64_t* r_off, const int32_t* r_len,
                                   const int32_t* mask_len, const ssw_scoring* scoring);
int ssw_align_batch_multi_packed(const int* devices, int n_devices, int32_t n_pairs, const uint8_t* packed, int64_t n_bases,
                                 const int64_t* q_off, const int32_t* q_len,
                                 const int64_t* r_off, const int32_t* r_len,
                                 const int32_t* mask_len, const ssw_scoring* scoring,
                                 ssw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used);
void ssw_pack_dna4(const int8_t* codes, int64_t n, uint8_t* packed);

/* ASCII -> {0..4} on the device-facing side of the boundary: vectorised replacement for the
 * per-base Python loop of ssw_wrap.py:234-252 (A C G T N, either case; anything else -> 4). */
void ssw_encode_dna(const char* ascii, int64_t len, int8_t* codes);

/* Diagnostics */
const char* ssw_cuda_last_error(void);
/* Batched unit-cost global edit distance.  Replaces
 * CIRI_long/utils.py:153-159 `distance(x, y)` -- Levenshtein.distance / edlib.align(...)['editDistance'],
 * the same quantity -- as called per pair from collapse.py:156-158 and collapse.py:466-473.
 * seqs: raw bytes (symbols are compared for equality; at most 16 distinct byte values per batch),
 * pair p = seqs[x_off[p] .. +x_len[p]) vs seqs[y_off[p] .. +y_len[p]); out[p] = distance.
 * Returns SSW_OK or an SSW_ERR_* code (ssw_cuda_last_error() says why). */
int ssw_cuda_edit_distance_batch(int device, int32_t n_pairs, const uint8_t* seqs, int64_t seqs_len,
                                 const int64_t* x_off, const int32_t* x_len,
                                 const int64_t* y_off, const int32_t* y_len, int32_t* out);
int ssw_cuda_device_count(void);
/* Device memory is cached in a pool owned by this library between batches; this returns it to the driver. */
int ssw_cuda_trim_pools(void);
/* Measured issue rate of a dependency-free VIADDMNMX.S16x2 stream on `device`, in 32-bit
 * lane-instructions per second (the denominator of the DPX roofline). */
int ssw_cuda_dpx_peak(int device, double* lane_instr_per_s, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* SSW_CUDA_H */
