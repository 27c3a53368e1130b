// ssw_tband_core.h -- lane program of the throughput CIGAR kernel (ssw_tband.cu): banded affine DP with
// 4-bit direction codes + traceback, ONE PAIR PER LANE, two query rows per register (s16x2).
//
// Replaces banded_sw (reference ssw.c:548-735) for large batches.  The reference fills the band in scalar
// int32 code; the warp-per-pair kernel of ssw_band.cu keeps one pair per warp (right for a handful of
// pairs, 2.8 ALU instructions per cell).  Here every lane owns a whole pair, so there are no shuffles, no
// pipeline fill/drain and no idle lanes on narrow bands:
//
//   * cells are addressed by band diagonal kk = j - i + w.  A lane walks its pair two rows at a time: the low
//     16-bit half of every register holds row i0 = 2*rho at diagonal t, the high half row i0+1 at diagonal
//     t-2 (same anti-diagonal, so the two cells are independent).  The upper row's results reach the lower
//     row through the registers of the previous two steps; only every second row goes through memory;
//   * the 32 lanes of a warp run their pairs in lock step (same rho, same t), so every per-lane array in
//     shared memory is indexed [position][lane]: conflict free by construction.  Pairs are sorted by (band
//     width, rows) before the launch, so lock step costs little (ssw_tband.cu);
//   * all values carry a bias B >= gap_open + gap_extend ("zero" is B): every subtraction of a gap penalty is
//     a plain 32-bit subtract that cannot borrow between the halves, the floors max(E,0) / max(F,0) are a
//     max with B, and each max of the recurrence is one max.s16x2 whose two predicate outputs are OR-ed
//     straight into the direction word (no SEL/LOP chains);
//   * 4 bits per cell: bit0 vertical gap opened (ssw.c:607-611 code 3 vs 2), bit1 horizontal gap opened
//     (ssw.c:613-616 code 5 vs 4), bit2 a gap beats the diagonal (ssw.c:624 strict), bit3 the horizontal gap
//     wins the tie against the vertical one (ssw.c:626).  One 32-bit word per 4 steps and lane, written as a
//     full 128-byte line per warp;
//   * cells outside band or rectangle read as zero (ssw.c:592-596) -- applied by masks only in the first and
//     last blocks of a row, where some lane can be outside (the "masked" body); the zeroed vertical neighbour
//     of the last column in the first w+1 rows (ssw.c:595-596, see oracle/ssw_oracle.c) lives there too.
//
// The same source compiles for the host (tools/tband_host_check.cpp runs 32 emulated lanes against the
// oracle); device-only pieces are in ssw_tband.cu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TB_HD __device__ __forceinline__
#else
#define TB_HD inline
#endif

namespace sswt {

constexpr int TB_BLOCK = 4;                     // steps per direction word
constexpr int TB_MAX_STEPS = 128;               // widest instance: 2*bw + 3 <= 128 (wider bands fill a whole warp: ssw_band.cu)
constexpr int TB_SCORE_LIMIT = 32000;           // score1 + bias must stay below the s16 range

// ---- packed primitives ---------------------------------------------------------------------------------
TB_HD unsigned tb_prmt(unsigned a, unsigned b, unsigned sel)
{
#if defined(__CUDACC__)
    unsigned v; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(a), "r"(b), "r"(sel)); return v;
#else
    const unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned v = 0;
    for (int k = 0; k < 4; ++k) v |= (unsigned)((src >> (8 * ((sel >> (4 * k)) & 7))) & 0xff) << (8 * k);
    return v;
#endif
}

TB_HD unsigned tb_max2(unsigned a, unsigned b)
{
#if defined(__CUDACC__)
    unsigned v; asm("max.s16x2 %0, %1, %2;" : "=r"(v) : "r"(a), "r"(b)); return v;
#else
    const short al = (short)(a & 0xffff), ah = (short)(a >> 16), bl = (short)(b & 0xffff), bh = (short)(b >> 16);
    return (unsigned)(unsigned short)(al > bl ? al : bl) | ((unsigned)(unsigned short)(ah > bh ? ah : bh) << 16);
#endif
}

// r = max.s16x2(a, b); per half: acc |= IMM if the maximum is NOT a  (a loses strictly:  b > a)
template <unsigned IMM_LO, unsigned IMM_HI>
TB_HD unsigned tb_max2_b_wins(unsigned a, unsigned b, unsigned& acc, const unsigned one)
{
#if defined(__CUDACC__)
    unsigned r;
    // (the high half's bit is added as a multiply-add by a register that holds 1: an IMAD on the FMA pipe instead of a
    // second add on the ALU pipe, which the maxima already keep busy)
    asm("{\n\t.reg .pred pl, ph;\n\t.reg .s16 r0, r1, a0, a1;\n\t"
        "max.s16x2 %0, %2, %3;\n\t"
        "mov.b32 {r0, r1}, %0;\n\t"
        "mov.b32 {a0, a1}, %2;\n\t"
        "setp.eq.s16 pl, r0, a0;\n\t"
        "setp.eq.s16 ph, r1, a1;\n\t"
        "@!pl or.b32 %1, %1, %4;\n\t"
        "@!ph mad.lo.u32 %1, %6, %5, %1;\n\t}"
        : "=&r"(r), "+r"(acc) : "r"(a), "r"(b), "n"(IMM_LO), "n"(IMM_HI), "r"(one));
    return r;
#else
    const unsigned r = tb_max2(a, b);
    (void)one;
    if ((r & 0xffff) != (a & 0xffff)) acc |= IMM_LO;
    if ((r >> 16) != (a >> 16)) acc |= IMM_HI;
    return r;
#endif
}

// r = max.s16x2(a, b); per half: acc |= IMM if the maximum IS a  (a >= b)
template <unsigned IMM_LO, unsigned IMM_HI>
TB_HD unsigned tb_max2_a_wins(unsigned a, unsigned b, unsigned& acc, const unsigned one)
{
#if defined(__CUDACC__)
    unsigned r;
    asm("{\n\t.reg .pred pl, ph;\n\t.reg .s16 r0, r1, a0, a1;\n\t"
        "max.s16x2 %0, %2, %3;\n\t"
        "mov.b32 {r0, r1}, %0;\n\t"
        "mov.b32 {a0, a1}, %2;\n\t"
        "setp.eq.s16 pl, r0, a0;\n\t"
        "setp.eq.s16 ph, r1, a1;\n\t"
        "@pl or.b32 %1, %1, %4;\n\t"
        "@ph mad.lo.u32 %1, %6, %5, %1;\n\t}"
        : "=&r"(r), "+r"(acc) : "r"(a), "r"(b), "n"(IMM_LO), "n"(IMM_HI), "r"(one));
    return r;
#else
    const unsigned r = tb_max2(a, b);
    (void)one;
    if ((r & 0xffff) == (a & 0xffff)) acc |= IMM_LO;
    if ((r >> 16) == (a >> 16)) acc |= IMM_HI;
    return r;
#endif
}

// ---- per-lane score tables: tab[c] = s(c, read[i0]) sign-extended, tab[5+c] = s(c, read[i0+1]) << 16 ------
// On the device the table is addressed through the 32-bit shared window, so that the address of an entry is
// one multiply-add (FMA pipe) and the second table is an immediate offset.
#if defined(__CUDACC__)
typedef unsigned TbAddr;
__device__ __forceinline__ TbAddr tb_tab_addr(const unsigned tabS, const unsigned c)
{ unsigned v; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(c), "r"(128u), "r"(tabS)); return v; }
__device__ __forceinline__ unsigned tb_tab_lo(const TbAddr a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned tb_tab_hi(const TbAddr a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1+640];" : "=r"(v) : "r"(a)); return v; }
typedef unsigned TbTab;
#else
typedef const int* TbAddr;
inline TbAddr tb_tab_addr(const int* tab, const unsigned c) { return tab + c * 32; }
inline unsigned tb_tab_lo(const TbAddr a) { return (unsigned)a[0]; }
inline unsigned tb_tab_hi(const TbAddr a) { return (unsigned)a[5 * 32]; }
typedef const int* TbTab;
#endif

// ---- per-lane state -------------------------------------------------------------------------------------
// Shared-memory arrays are passed as pointers already offset to the lane; consecutive positions are
// TB_LANES elements apart.
constexpr int TB_LANES = 32;

struct TbJob {                       // one pass of one pair (ssw.c:571-632: one iteration of the do-loop)
    const int8_t* ref;               // trimmed rectangle: ref[0 .. refLen)
    const int8_t* read;              //                    read[0 .. readLen)
    int32_t refLen, readLen;
    int32_t bw;                      // band half-width of this pass
    int32_t score;                   // score1: the pass is final when max >= score or bw >= readLen
    int32_t maxIn;                   // running maximum carried from narrower passes
};

struct TbRow {                       // state of the current row pair
    int32_t aL, bL, aH, bH;          // valid step ranges [a, b] of the low / high row (time index t)
    int32_t tqL, tqH;                // step whose vertical neighbour is zeroed (ssw.c:595-596), -1 = none
    unsigned Hl, Fl, Eprev, Hd;      // H / F of the previous step (left neighbours), E of it, diagonal
    TbAddr cPrevAddr;                // table address of the previous reference base (high row's score)
};

TB_HD int tb_steps(int bw) { return 2 * bw + 3; }                           // 2w+1 diagonals + 2 of skew
TB_HD int tb_blocks(int bw) { return (tb_steps(bw) + TB_BLOCK - 1) / TB_BLOCK; }

// valid diagonals of row i: j = i + kk - w inside [max(0,i-w), min(refLen-1,i+w)]  (ssw.c:592-594)
TB_HD void tb_row_range(int i, int readLen, int refLen, int w, int& a, int& b)
{
    if (i >= readLen) { a = 1; b = 0; return; }
    a = w - i > 0 ? w - i : 0;
    const int hi = refLen - 1 - i + w;
    b = hi < 2 * w ? hi : 2 * w;
}
// ssw.c:595-596: in rows 1..w+1 whose band already reaches the last reference column, the vertical
// neighbour of the cell on that column reads as zero.  Returns that cell's diagonal or -1.
TB_HD int tb_row_quirk(int i, int readLen, int refLen, int w)
{
    if (i < 1 || i >= readLen || i - 1 - w > 0 || i - 1 + w < refLen - 1) return -1;
    return refLen - 1 - i + w;
}

// One step.  S: H|E<<16 of the row above the pair, indexed by diagonal (in place: this step reads t+1 and
// writes t-2).  ringB: reference codes by band position, bytes.  MODE says what can be outside band or
// rectangle in this block (the driver picks it per block for the whole warp):
//   TB_PLAIN  nothing: both rows inside for every lane
//   TB_HEAD   block 0 of a row pair whose band starts at diagonal 0: only the high row's two steps of skew
//   TB_TAIL   the last blocks: a step is valid iff t <= b (the row has begun in every lane)
//   TB_ANY    anything, including the zeroed vertical neighbour of ssw.c:595-596
enum { TB_PLAIN = 0, TB_ANY = 1, TB_HEAD = 2, TB_TAIL = 3 };
template <int P, int MODE>
TB_HD void tb_step(TbRow& R, const int t, unsigned* S, const unsigned char* ringPos, const TbTab tab,
                   const unsigned B2, const unsigned GO2, const unsigned GE2, const unsigned one, unsigned& dirw, unsigned& maxv2)
{
    unsigned w = S[(t + 1) * TB_LANES];
    unsigned Hprev = R.Hl, Eprev = R.Eprev;
    if (MODE == TB_ANY) {
        if (t == R.tqL) w = B2;                                   // low row: zeroed vertical neighbour
        if (t == R.tqH) { Hprev = (Hprev & 0xffff0000u) | (B2 & 0xffffu); Eprev = (Eprev & 0xffff0000u) | (B2 & 0xffffu); }
    }
    // vertical neighbours: low row from memory, high row = the low row's cell of the previous step
    const unsigned Hup = tb_prmt(w, Hprev, 0x5410u);
    const unsigned Eup = tb_prmt(w, Eprev, 0x5432u);
    const unsigned eopen = Hup - GO2, eext = Eup - GE2;
    unsigned E = tb_max2_b_wins<1u << (8 * P), 1u << (8 * P + 4)>(eext, eopen, dirw, one);        // open only if strictly greater
    const unsigned fopen = R.Hl - GO2, fext = R.Fl - GE2;
    unsigned F = tb_max2_b_wins<2u << (8 * P), 2u << (8 * P + 4)>(fext, fopen, dirw, one);
    const unsigned f1 = tb_max2(F, B2);
    const unsigned gb = tb_max2_a_wins<8u << (8 * P), 8u << (8 * P + 4)>(f1, E, dirw, one);       // F wins ties against E
    // substitution scores: low row against ref[j], high row against ref[j-1]
    const TbAddr cAddr = tb_tab_addr(tab, (unsigned)ringPos[P * TB_LANES]);
    const unsigned dg = R.Hd + tb_tab_lo(cAddr) + tb_tab_hi(R.cPrevAddr);
    R.cPrevAddr = cAddr;
    unsigned H = tb_max2_b_wins<4u << (8 * P), 4u << (8 * P + 4)>(dg, gb, dirw, one);             // diagonal wins ties
    bool hiValid = true;
    if (MODE == TB_ANY || MODE == TB_TAIL) {
        bool loValid;
        if (MODE == TB_ANY) {
            loValid = (unsigned)(t - R.aL) <= (unsigned)(R.bL - R.aL) && R.bL >= R.aL;
            hiValid = (unsigned)(t - R.aH) <= (unsigned)(R.bH - R.aH) && R.bH >= R.aH;
        } else { loValid = t <= R.bL; hiValid = t <= R.bH; }
        const unsigned m = (loValid ? 0xffffu : 0u) | (hiValid ? 0xffff0000u : 0u);
        H = (H & m) | (B2 & ~m); E = (E & m) | (B2 & ~m); F = (F & m) | (B2 & ~m);
    }
    if (MODE == TB_HEAD && P < 2) {
        hiValid = false;
        H = (H & 0xffffu) | (B2 & 0xffff0000u); E = (E & 0xffffu) | (B2 & 0xffff0000u); F = (F & 0xffffu) | (B2 & 0xffff0000u);
    }
    maxv2 = tb_max2(maxv2, H);
    // the high row is the one the next row pair reads back
    if (hiValid) S[(t - 2) * TB_LANES] = tb_prmt(H, E, 0x7632u);
    R.Hd = Hup;
    R.Hl = H; R.Eprev = E; R.Fl = F;
}

// One block of TB_BLOCK steps; returns the direction word.
template <int MODE>
TB_HD unsigned tb_block(TbRow& R, const int t0, unsigned* S, const unsigned char* ringPos, const TbTab tab,
                        const unsigned B2, const unsigned GO2, const unsigned GE2, const unsigned one, unsigned& maxv2)
{
    unsigned dirw = 0;
    tb_step<0, MODE>(R, t0 + 0, S, ringPos, tab, B2, GO2, GE2, one, dirw, maxv2);
    tb_step<1, MODE>(R, t0 + 1, S, ringPos, tab, B2, GO2, GE2, one, dirw, maxv2);
    tb_step<2, MODE>(R, t0 + 2, S, ringPos, tab, B2, GO2, GE2, one, dirw, maxv2);
    tb_step<3, MODE>(R, t0 + 3, S, ringPos, tab, B2, GO2, GE2, one, dirw, maxv2);
    return dirw;
}

// Which blocks of a row pair need which body.  Per lane: the first step from which both rows are inside
// (lo) and the last one (hi); `simple` = the band starts at diagonal 0 and nothing but the skew is outside at
// the front.  The driver reduces (max lo, min hi, all simple) over the lanes that still have rows.
struct TbRowPlan { int lo, hi; bool simple; };
TB_HD TbRowPlan tb_row_plan(const TbRow& R, const int NB)
{
    TbRowPlan p;
    p.lo = R.aL > R.aH ? R.aL : R.aH;
    p.hi = R.bL < R.bH ? R.bL : R.bH;
    p.simple = R.aL == 0 && R.aH == 2 && R.tqL < 0 && R.tqH < 0 && p.hi >= p.lo;
    if (R.tqL >= 0 || R.tqH >= 0 || p.hi < p.lo) { p.lo = 4 * NB; p.hi = -1; }
    return p;
}

// Start of a row pair: ranges, score tables, neighbour state.  matS: the 5x5 matrix as ints; r0, r1: the read
// bases of the two rows (already clamped to 0..4; the driver loads them one row pair ahead).
TB_HD void tb_row_begin(TbRow& R, const TbJob& J, const int rho, const unsigned* S, int* tab, const TbTab tabRef, const int* matS,
                        const unsigned B2, const unsigned r0, const unsigned r1)
{
    const int i0 = 2 * rho, i1 = i0 + 1;
    int a, b;
    tb_row_range(i0, J.readLen, J.refLen, J.bw, a, b); R.aL = a; R.bL = b;
    tb_row_range(i1, J.readLen, J.refLen, J.bw, a, b); R.aH = a + 2; R.bH = b + 2;
    const int qL = tb_row_quirk(i0, J.readLen, J.refLen, J.bw), qH = tb_row_quirk(i1, J.readLen, J.refLen, J.bw);
    R.tqL = qL; R.tqH = qH < 0 ? -1 : qH + 2;
    for (int c = 0; c < 5; ++c) {
        tab[c * TB_LANES] = matS[c * 5 + r0];
        tab[(5 + c) * TB_LANES] = matS[c * 5 + r1] * 65536;
    }
    R.Hl = B2; R.Fl = B2; R.Eprev = B2;
    R.Hd = (S[0] & 0xffffu) | (B2 & 0xffff0000u);                 // diagonal of (i0, 0) is (i0-1, 0); high row starts outside
    R.cPrevAddr = tb_tab_addr(tabRef, 0);                         // high row's first two steps are outside the band
}
TB_HD unsigned tb_read_code(const TbJob& J, int i)
{
    const int last = J.readLen - 1;
    const unsigned c = (unsigned char)J.read[i < last ? i : last];
    return c > 4u ? 4u : c;                                       // codes above 4 are never produced by the encoders; they count as N
}
TB_HD unsigned tb_ref_code(const TbJob& J, int col)
{
    col = col < 0 ? 0 : (col > J.refLen - 1 ? J.refLen - 1 : col);
    const unsigned c = (unsigned char)J.ref[col];
    return c > 4u ? 4u : c;
}

// ---- traceback (ssw.c:636-727) ---------------------------------------------------------------------------
// dirs: this lane's direction words, word (rho, block) at dirs[(rho * NB + block) * TB_LANES].
// stage: this lane's op staging (traceback order), stage[k * TB_LANES].  Returns the number of ops or -1 (band
// left: the reference's "Trace back error") or -2 (staging full).
TB_HD int tb_traceback(const TbJob& J, const unsigned* dirs, const int NB, unsigned* stage, const int stageCap)
{
    const int w = J.bw;
    int i = J.readLen - 1, j = J.refLen - 1, state = 2;
    int op = 0, prevOp = 0, run = 0, nOps = 0;                    // 0 M, 1 I, 2 D
    while (i > 0) {
        const int beg = i - w > 0 ? i - w : 0;
        const int end = i + w < J.refLen - 1 ? i + w : J.refLen - 1;
        if (j < beg || j > end) return -1;
        const int kk = j - i + w;
        const int half = i & 1, t = kk + 2 * half;
        const unsigned word = dirs[((i >> 1) * NB + (t >> 2)) * TB_LANES];
        const unsigned n = (word >> (8 * (t & 3) + 4 * half)) & 0xfu;
        int code;
        if (state == 2) code = !(n & 4u) ? 1 : ((n & 8u) ? 4 + (int)((n >> 1) & 1u) : 2 + (int)(n & 1u));
        else if (state == 0) code = 2 + (int)(n & 1u);
        else code = 4 + (int)((n >> 1) & 1u);
        switch (code) {
            case 1: --i; --j; state = 2; op = 0; break;
            case 2: --i; state = 0; op = 1; break;
            case 3: --i; state = 2; op = 1; break;
            case 4: --j; state = 1; op = 2; break;
            default: --j; state = 2; op = 2; break;
        }
        if (op == prevOp) ++run;
        else {
            if (nOps + 3 >= stageCap) return -2;
            stage[nOps * TB_LANES] = ((unsigned)run << 4) | (unsigned)prevOp; ++nOps;
            prevOp = op; run = 1;
        }
    }
    if (nOps + 3 >= stageCap) return -2;
    if (op == 0) { stage[nOps * TB_LANES] = ((unsigned)(run + 1) << 4); ++nOps; }              // ssw.c:697-704
    else { stage[nOps * TB_LANES] = ((unsigned)run << 4) | (unsigned)op; ++nOps; stage[nOps * TB_LANES] = (1u << 4); ++nOps; }
    return nOps;
}

}  // namespace sswt
