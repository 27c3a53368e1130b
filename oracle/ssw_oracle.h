/* ssw_oracle.h -- TEST INFRASTRUCTURE ONLY (see ssw_oracle.c). */
#ifndef SSW_ORACLE_H
#define SSW_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_OK = 0, ORC_ERR_OVERFLOW = -1, ORC_ERR_NOPROFILE = -2, ORC_ERR_TRACEBACK = -3, ORC_ERR_CAPACITY = -4 };

/* result of one score pass: the reference's alignment_end[2] (ssw.c:67-71) flattened */
typedef struct { int32_t score, ref, read, score2, ref2; } orc_ends;

/* s_align (ssw.h:42-52) plus diagnostics (word = 1 if the 16-bit flavour produced the result) */
typedef struct {
    uint16_t score1, score2;
    int32_t ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    uint32_t* cigar;
    int32_t cigarLen;
    int32_t word;
    int32_t band_width;
} orc_result;

typedef struct {
    int32_t status, word, band_width;
    int32_t score1, score2, ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    int64_t cigar_off;
    int32_t cigar_len;
    int32_t _pad;
} orc_flat;

static inline uint32_t orc_to_cigar_int(uint32_t length, char op)
{
    uint32_t code = 0;                                  /* ssw.h:132-170 */
    switch (op) {
        case 'I': code = 1; break; case 'D': code = 2; break; case 'N': code = 3; break;
        case 'S': code = 4; break; case 'H': code = 5; break; case 'P': code = 6; break;
        case '=': code = 7; break; case 'X': code = 8; break; default: code = 0; break;
    }
    return (length << 4) | code;
}

void orc_score_pass(int word, const int8_t* ref, int ref_dir, int32_t refLen,
                    const int8_t* read, int32_t readLen, const int8_t* mat, int32_t n,
                    uint8_t gapO, uint8_t gapE, int32_t terminate, uint8_t bias, int32_t maskLen,
                    orc_ends* out);

int orc_band_cigar(const int8_t* ref, const int8_t* read, int32_t refLen, int32_t readLen,
                   int32_t score, uint32_t gapO, uint32_t gapE, int32_t band_width,
                   const int8_t* mat, int32_t n, uint32_t** cigar_out, int32_t* cigar_len_out,
                   int32_t* final_band_out);

int orc_align(const int8_t* read, int32_t readLen, const int8_t* ref, int32_t refLen,
              const int8_t* mat, int32_t n, int8_t score_size,
              uint8_t gapO, uint8_t gapE, uint8_t flag, uint16_t filters, int32_t filterd,
              int32_t maskLen, orc_result* r);
void orc_result_free(orc_result* r);
char orc_cigar_op(uint32_t c);
uint32_t orc_cigar_len(uint32_t c);

int orc_align_batch(int32_t n_pairs, const int8_t* seqs,
                    const int64_t* q_off, const int32_t* q_len,
                    const int64_t* r_off, const int32_t* r_len,
                    const int8_t* mat, int32_t n, uint8_t gapO, uint8_t gapE, uint8_t flag,
                    const int32_t* mask_len,
                    orc_flat* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used);

#ifdef __cplusplus
}
#endif
#endif
