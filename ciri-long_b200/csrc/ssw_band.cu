// ssw_band.cu -- banded affine DP + traceback -> CIGAR on the trimmed rectangle.
//
// Replaces banded_sw (reference ssw.c:548-735) as called from ssw_align (ssw.c:852-856).  The
// reference fills the band cell by cell in scalar int32 code; here one warp owns a pair:
//
//   * cells are addressed by their band diagonal kk = j - i + w (w = band half-width): the vertical
//     neighbour of (i, kk) is (i-1, kk+1), the diagonal neighbour is (i-1, kk), the horizontal one is
//     (i, kk-1);
//   * lane l owns a block of DPL = 2P consecutive diagonals and keeps H and the vertical-gap score E of
//     the previous row of those diagonals in registers.  Lanes run a skewed wavefront: in iteration r
//     lane l computes row r - l, left to right inside its block, so the horizontal-gap chain F is a plain
//     register chain, the left neighbour of the block comes from lane l-1 (one iteration old) and the
//     vertical neighbour of the block's last diagonal from lane l+1 (this iteration): two shuffles of
//     (H, gap) per row instead of a scan;
//   * one direction byte per cell (vertical source, horizontal source, H source), written as one
//     aligned DPL-byte store per lane and row into a per-warp scratch matrix with a padded row stride;
//     the band doubles while max < score1 exactly like ssw.c:631-632;
//   * the traceback (ssw.c:642-696) walks the matrix through a shared-memory window that the warp
//     refills 4 KB at a time; lane 0 runs the reference's state machine, run-length encodes, and the warp
//     writes the reversed ops to the output buffer.
// Bands wider than 1024 diagonals (never seen on the benchmark shapes) use a row-parallel variant in
// which F is an exclusive max-plus prefix scan over the lanes.
// Tie rules, the "stop at read row 0" rule and the zeroed vertical neighbour of the last column in the
// first w+1 rows (ssw.c:595-596) are reproduced; see oracle/ssw_oracle.c:orc_band_cigar.
#include <type_traits>
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

constexpr int NEG_INF = -(1 << 29);
constexpr int TB_WINDOW = 4096;          // bytes of shared memory per warp for the traceback window

// direction byte: bit0 vertical gap opened (code 3) / extended (2); bit1 horizontal gap opened (5) /
// extended (4); bits 2-3 H source: 0 diagonal, 1 vertical gap, 2 horizontal gap
__device__ __forceinline__ uint32_t cigar_pack(uint32_t len, int op) { return (len << 4) | (uint32_t)op; }

struct BandJob {
    const int8_t* ref;      // trimmed rectangle
    const int8_t* read;
    int refLen, readLen, bw;
    int go, ge;
    const int8_t* mat;      // 5x5 in constant bank
};

// ---- wavefront variant: 32 lanes x DPL diagonals, one row per lane and iteration -------------------
// stab: shared table, stab[rd] = the five substitution scores against read base rd packed as bytes
// (A C G T in .x, N in .y); a cell picks its score with one PRMT whose selector (base * 0x1111 + 0x8880:
// byte index + sign replication) travels with the diagonal instead of the base code.
// QUIRK = the band can be clipped by the reference end within the first w+1 rows (ssw.c:595-596 applies).
template <int DPL, bool QUIRK>
__device__ __forceinline__ int band_fill_wave(const int8_t* __restrict__ ref, const int8_t* __restrict__ read,
                                              const int refLen, const int readLen, const int bw, const int go, const int ge,
                                              const uint2* stab, unsigned char* dir, const int rowStride, int maxv)
{
    constexpr int BPL = DPL == 3 ? 4 : DPL;                    // direction bytes per lane and row (3 is padded)
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    const int kkBase = lane * DPL;
    int Hp[DPL], Ep[DPL];
    unsigned sel[DPL];
    {
        const int j0 = -lane + kkBase - bw;                     // column of my first diagonal in the row of iteration 0
#pragma unroll
        for (int d = 0; d < DPL; ++d) {
            Hp[d] = 0; Ep[d] = 0;
            const int j = j0 + d;
            unsigned c = 4;
            if (j >= 0 && j < refLen) { c = (unsigned)ref[j]; if (c > 4u) c = 4; }
            sel[d] = c * 0x1111u + 0x8880u;
        }
    }
    int lastH = 0, lastF = 0;                                   // H, F of my last diagonal in the row I just finished
    const int iters = readLen + 31;
    int i = -lane;                                              // my row in this iteration
    int jFirst = -lane + kkBase - bw;                           // column of my first diagonal in row i
    const int8_t* nextRef = ref + jFirst + DPL;                 // base entering my last diagonal in the next row
    unsigned char* drow = dir + (long long)i * rowStride + lane * BPL;
    // lane-constant part of cell validity: diagonal inside the band (0/1, applied by multiplication so
    // that it runs on the FMA pipe in the interior iterations)
    int vd[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) vd[d] = kkBase + d <= 2 * bw ? 1 : 0;

    // One iteration.  INTERIOR = every lane that owns an in-band diagonal is on a row whose band lies fully
    // inside the rectangle: validity is the lane constant vd[], the clipping rule of ssw.c:595-596 cannot apply.
    auto iteration = [&](auto interior) {
        constexpr bool INTERIOR = decltype(interior)::value;
        const bool rowOk = INTERIOR || (unsigned)i < (unsigned)readLen;
        // left neighbour of my block: lane-1's last cell of the same row (computed one iteration ago)
        int Hl = __shfl_up_sync(FULL, lastH, 1), Fl = __shfl_up_sync(FULL, lastF, 1);
        if (lane == 0) { Hl = 0; Fl = 0; }
        unsigned rd = 4;
        if (INTERIOR) { if (vd[0]) { rd = (unsigned)read[i]; if (rd > 4u) rd = 4; } }
        else if (rowOk) { rd = (unsigned)read[i]; if (rd > 4u) rd = 4; }
        const uint2 srow = stab[rd];
        int span = 0, jrel = 0, jq = 0;
        bool quirk = false;
        if (!INTERIOR) {
            // columns of this row inside band and rectangle: beg <= j <= end  <=>  (unsigned)(j - beg) <= span
            const int beg = i - bw > 0 ? i - bw : 0;
            const int end = i + bw < refLen - 1 ? i + bw : refLen - 1;
            const bool rowHas = rowOk && end >= beg;             // rows past the band's reach hold no cell
            span = rowHas ? end - beg : 0;
            jrel = rowHas ? jFirst - beg : -(1 << 30);           // (unsigned) of a negative never passes the test
            quirk = QUIRK && i >= 1 && i - 1 - bw <= 0 && i - 1 + bw >= refLen - 1;
            jq = refLen - 1 - jFirst;                            // diagonal slot that sits on the last column
        }
        unsigned dirw[(DPL + 3) / 4];
#pragma unroll
        for (int w = 0; w < (DPL + 3) / 4; ++w) dirw[w] = 0;
        int upH = 0, upE = 0;                                   // (i-1, first diagonal of lane+1), known after d == 0
#pragma unroll
        for (int d = 0; d < DPL; ++d) {
            int Hu, Eu;
            if (d + 1 < DPL) { Hu = Hp[d + 1]; Eu = Ep[d + 1]; }
            else { Hu = upH; Eu = upE; }
            if (QUIRK && !INTERIOR) { if (quirk && d == jq) { Hu = 0; Eu = 0; } }
            const int s = (int)prmt_raw(srow.x, srow.y, sel[d]);
            // each max also delivers the comparison the direction code needs (VIMNMX with predicate output)
            const int eopen = Hu - go, eext = Eu - ge;
            bool eExtWins, fExtWins, fWins, diagWins;
            const int E = __vibmax_s32(eext, eopen, &eExtWins);            // eExtWins = (eext >= eopen): open only if strictly greater
            const int fopen = Hl - go, fext = Fl - ge;
            const int F = __vibmax_s32(fext, fopen, &fExtWins);
            const int e1 = E > 0 ? E : 0, f1 = F > 0 ? F : 0;
            const int gapbest = __vibmax_s32(f1, e1, &fWins);               // fWins = (f1 >= e1): E only if strictly greater
            const int dg = Hp[d] + s;
            int H = __vibmax_s32(dg, gapbest, &diagWins);                   // diagWins = (dg >= gapbest)
            unsigned code = (eExtWins ? 0u : 1u) | (fExtWins ? 0u : 2u);
            if (!diagWins) code |= fWins ? 8u : 4u;
            int Eo = E, Fo = F;
            if (INTERIOR) { H *= vd[d]; Eo *= vd[d]; Fo *= vd[d]; }
            else if (!((unsigned)(jrel + d) <= (unsigned)span)) { H = 0; Eo = 0; Fo = 0; }
            maxv = H > maxv ? H : maxv;
            dirw[d >> 2] |= code << (8 * (d & 3));
            Hp[d] = H; Ep[d] = Eo;
            Hl = H; Fl = Fo;
            if (d == 0) {
                // vertical neighbour of my last diagonal: lane+1's first cell of row i-1, which lane+1
                // has just computed in this iteration
                upH = __shfl_down_sync(FULL, H, 1);
                upE = __shfl_down_sync(FULL, Eo, 1);
                if (lane == 31) { upH = 0; upE = 0; }
            }
        }
        lastH = Hl; lastF = Fl;
        if (INTERIOR ? vd[0] != 0 : rowOk) {
            if (DPL == 2) *reinterpret_cast<unsigned short*>(drow) = (unsigned short)dirw[0];
            else if (DPL == 3 || DPL == 4) *reinterpret_cast<unsigned*>(drow) = dirw[0];
            else if (DPL == 8) *reinterpret_cast<uint2*>(drow) = make_uint2(dirw[0], dirw[1]);
            else {
#pragma unroll
                for (int w = 0; w < DPL / 16; ++w)
                    reinterpret_cast<uint4*>(drow)[w] = make_uint4(dirw[4 * w], dirw[4 * w + 1], dirw[4 * w + 2], dirw[4 * w + 3]);
            }
        }
        // next row: every diagonal moves one reference base to the right
#pragma unroll
        for (int d = 0; d + 1 < DPL; ++d) sel[d] = sel[d + 1];
        {
            const int j = jFirst + DPL;
            unsigned c = 4;
            if ((unsigned)j < (unsigned)refLen) { c = (unsigned)*nextRef; if (c > 4u) c = 4; }
            sel[DPL - 1] = c * 0x1111u + 0x8880u;
        }
        ++i; ++jFirst; ++nextRef; drow += rowStride;
    };

    // iterations [rA, rB) are interior for every lane that owns in-band diagonals (lanes 0 .. lmax)
    const int lmax = (2 * bw) / DPL;
    int rA = bw + lmax, rB = (readLen - 1 < refLen - 1 - bw ? readLen - 1 : refLen - 1 - bw) + 1;
    if (rA > iters) rA = iters;
    if (rB < rA) rB = rA;
    if (rB > iters) rB = iters;
    int r = 0;
    for (; r < rA; ++r) iteration(std::false_type{});
    for (; r < rB; ++r) iteration(std::true_type{});
    for (; r < iters; ++r) iteration(std::false_type{});
    return __reduce_max_sync(FULL, maxv);
}

// ---- row-parallel variant for very wide bands: F by max-plus prefix scan ---------------------------
__device__ int band_fill_scan(const BandJob& jb, int* Hrow, unsigned char* dir, int rowStride, int maxv)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    const int bw = jb.bw, go = jb.go, ge = jb.ge, refLen = jb.refLen, readLen = jb.readLen;
    int* Erow = Hrow + (2 * bw + 4);
    for (int k = lane; k < 2 * bw + 4; k += 32) { Hrow[k] = 0; Erow[k] = 0; }
    __syncwarp();
    for (int i = 0; i < readLen; ++i) {
        const int beg = i - bw > 0 ? i - bw : 0;
        const int end = i + bw < refLen - 1 ? i + bw : refLen - 1;
        if (beg > end) break;
        const int W = end - beg + 1;
        const int kk0 = beg - i + bw;
        int rd = jb.read[i]; if ((unsigned)rd > 4u) rd = 4;
        const bool quirk = i >= 1 && i - 1 - bw <= 0 && i - 1 + bw >= refLen - 1;
        unsigned char* drow = dir + (size_t)i * rowStride;
        int carryV = NEG_INF, carryH = 0, carryF = 0;
        for (int t0 = 0; t0 < W; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < W;
            const int kk = kk0 + t;
            const int j = beg + t;
            int Hd = 0, Hu = 0, Eu = 0, s = 0;
            if (act) {
                Hd = Hrow[kk]; Hu = Hrow[kk + 1]; Eu = Erow[kk + 1];
                if (quirk && j == refLen - 1) { Hu = 0; Eu = 0; }
                int rf = jb.ref[j]; if ((unsigned)rf > 4u) rf = 4;
                s = jb.mat[rf * 5 + rd];
            }
            __syncwarp();
            const int eopen = Hu - go, eext = Eu - ge;
            const int E = eopen > eext ? eopen : eext;
            const int de = eopen > eext ? 1 : 0;
            const int e1 = E > 0 ? E : 0;
            const int dg = Hd + s;
            const int A = e1 > dg ? e1 : dg;
            int inc = act ? A + t * ge : NEG_INF;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d && o > inc) inc = o;
            }
            int exc = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) exc = NEG_INF;
            if (carryV > exc) exc = carryV;
            int F = -ge * (t + 1);
            if (exc > NEG_INF) { const int f2 = exc - go - (t - 1) * ge; if (f2 > F) F = f2; }
            const int H = A > F ? A : F;
            int Hl = __shfl_up_sync(FULL, H, 1), Fl = __shfl_up_sync(FULL, F, 1);
            if (lane == 0) { Hl = carryH; Fl = carryF; }
            const int df = (Hl - go > Fl - ge) ? 2 : 0;
            const int f1 = F > 0 ? F : 0;
            const int gapbest = e1 > f1 ? e1 : f1;
            int dh = 0;
            if (gapbest > dg) dh = e1 > f1 ? 4 : 8;
            if (act) {
                Hrow[kk] = H; Erow[kk] = E;
                drow[kk] = (unsigned char)(de | df | dh);
                if (H > maxv) maxv = H;
            }
            const int top = __shfl_sync(FULL, inc, 31);
            carryV = top > carryV ? top : carryV;
            carryH = __shfl_sync(FULL, H, 31);
            carryF = __shfl_sync(FULL, F, 31);
            __syncwarp();
        }
    }
    return __reduce_max_sync(FULL, maxv);
}

// CLASS 0: bands of up to 128 diagonals (the common case, 64 registers); CLASS 1: up to 256 (8 diagonals per
// lane, 128 registers); CLASS 2: everything wider (255 registers).  A pair whose band grows past its class
// is handed to the next instance (its own launch) together with the band width reached and the
// running maximum, which is all the doubling loop carries from one width to the next (ssw.c:571-632).
template <int CLASS>
__device__ void band_pair(const BandArgs& a, const int pair, unsigned char* ws, unsigned char* win, const uint2* stab)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = lane_id();
    PairRec* rec = a.b.rec + pair;
    const int refBeg = rec->ref_begin1, readBeg = rec->read_begin1;
    const int score = rec->score1;

    uint32_t* stage = reinterpret_cast<uint32_t*>(ws);                         // cigar ops, traceback order
    unsigned char* area = ws + (((size_t)a.cigar_stage_cap * 4 + 15) & ~(size_t)15);   // row buffers + direction matrix
    int nOps = 0;
    int status = 0;

    if (refBeg < 0) {
        // score 0 in the byte flavour: the reference reads ref[-1] (undefined) on a 1x1 rectangle; the
        // traceback loop never runs and the CIGAR is "1M" (oracle/ssw_oracle.c documents the same rule).
        if (lane == 0) stage[0] = cigar_pack(1, 0);
        nOps = 1;
    } else {
        BandJob jb;
        jb.refLen = rec->ref_end1 - refBeg + 1;
        jb.readLen = rec->read_end1 - readBeg + 1;
        jb.ref = a.b.seqs + a.b.r_off[pair] + refBeg;
        jb.read = a.b.seqs + a.b.q_off[pair] + readBeg;
        jb.go = a.sc.go; jb.ge = a.sc.ge; jb.mat = a.sc.mat;
        const int readLen = jb.readLen, refLen = jb.refLen;
        int bw = refLen - readLen; if (bw < 0) bw = -bw; bw += 1;
        int maxv = 0, rowStride = 0, dpl = 0;
        if (CLASS > 0) { bw = -rec->cigar_len; maxv = (int)rec->cigar_off; }
        unsigned char* dir = nullptr;
        for (;;) {
            const int Wd = 2 * bw + 1;
            jb.bw = bw;
            // diagonals per lane (0 = scan variant); at least 2, so that a lane's first cell has its vertical
            // neighbour in the lane's own registers and the shuffle can sit between its first and last cell
            int P = 0;
            if (Wd <= 64) P = 2; else if (Wd <= 96) P = 3; else if (Wd <= 128) P = 4;
            else if (Wd <= 256) P = 8; else if (Wd <= 512) P = 16; else if (Wd <= 1024) P = 32;
            if ((CLASS == 0 && (P == 0 || P > 4)) || (CLASS == 1 && (P == 0 || P > 8))) {
                if (lane == 0) {
                    rec->cigar_len = -bw; rec->cigar_off = maxv;
                    if (CLASS == 0 && P == 8) a.next_idx[atomicAdd(a.next_count, 1)] = pair;
                    else a.next2_idx[atomicAdd(a.next2_count, 1)] = pair;
                }
                return;
            }
            dpl = P;
            rowStride = P ? 32 * (P == 3 ? 4 : P) : ((Wd + 15) & ~15);
            const long long rowBufBytes = P ? 0 : (((long long)(2 * bw + 4) * 8 + 15) & ~15LL);
            const long long need = rowBufBytes + (long long)rowStride * readLen;
            if (need > a.dir_bytes) { status = PS_BAND_SCRATCH; break; }
            dir = area + rowBufBytes;
            // ssw.c:595-596 can only matter if the band reaches the last column within the first w+1 rows
            const bool q = 2 * bw + 1 >= refLen - 1;
#define SSW_WAVE(PP) (q ? band_fill_wave<PP, true>(jb.ref, jb.read, refLen, readLen, bw, jb.go, jb.ge, stab, dir, rowStride, maxv) \
                        : band_fill_wave<PP, false>(jb.ref, jb.read, refLen, readLen, bw, jb.go, jb.ge, stab, dir, rowStride, maxv))
            if (CLASS == 0) {
                switch (P) {
                    case 2: maxv = SSW_WAVE(2); break;
                    case 3: maxv = SSW_WAVE(3); break;
                    default: maxv = SSW_WAVE(4); break;
                }
            } else if (CLASS == 1) {
                switch (P) {                                     // (a pair can arrive with its band already at 8 per lane only)
                    case 2: maxv = SSW_WAVE(2); break;
                    case 3: maxv = SSW_WAVE(3); break;
                    case 4: maxv = SSW_WAVE(4); break;
                    default: maxv = SSW_WAVE(8); break;
                }
            } else {
                switch (P) {
                    case 2: maxv = SSW_WAVE(2); break;
                    case 3: maxv = SSW_WAVE(3); break;
                    case 4: maxv = SSW_WAVE(4); break;
                    case 8: maxv = SSW_WAVE(8); break;
                    case 16: maxv = SSW_WAVE(16); break;
                    case 32: maxv = SSW_WAVE(32); break;
                    default: maxv = band_fill_scan(jb, reinterpret_cast<int*>(area), dir, rowStride, maxv); break;
                }
            }
#undef SSW_WAVE
            bw *= 2;
            if (!(maxv < score && bw < 2 * readLen)) break;
        }
        bw /= 2;

        if (!status) {
            __syncwarp();
            // traceback: start in state H at the bottom-right corner, stop at read row 0 (ssw.c:636-696)
            int i = readLen - 1, j = refLen - 1, state = 2;
            int op = 0, prevOp = 0, run = 0;                                   // 0 M, 1 I, 2 D
            int rowsPerWin = TB_WINDOW / rowStride; if (rowsPerWin < 1) rowsPerWin = 1;
            const bool windowed = rowStride <= TB_WINDOW;
            while (i > 0 && !status) {
                // rows (lo, i] go to shared memory, then lane 0 walks until it leaves them
                const int hi = i;
                int lo = hi - rowsPerWin; if (lo < 0) lo = 0;                  // rows lo+1 .. hi
                if (windowed) {
                    const uint4* src = reinterpret_cast<const uint4*>(dir + (size_t)(lo + 1) * rowStride);
                    const int n16 = (hi - lo) * rowStride / 16;
                    for (int k = lane; k < n16; k += 32) reinterpret_cast<uint4*>(win)[k] = src[k];
                    __syncwarp();
                }
                if (lane == 0) {
                    while (i > lo) {
                        const int beg = i - bw > 0 ? i - bw : 0;
                        const int end = i + bw < refLen - 1 ? i + bw : refLen - 1;
                        if (j < beg || j > end) { status = PS_TRACEBACK_ERR; break; }
                        int kk = j - i + bw;
                        if (dpl == 3) kk += kk / 3;                            // three diagonals per lane sit in four bytes
                        const int d = windowed ? win[(i - lo - 1) * rowStride + kk] : dir[(size_t)i * rowStride + kk];
                        int code;
                        if (state == 2) { const int dh = d >> 2; code = dh == 0 ? 1 : (dh == 1 ? 2 + (d & 1) : 4 + ((d >> 1) & 1)); }
                        else if (state == 0) code = 2 + (d & 1);
                        else code = 4 + ((d >> 1) & 1);
                        switch (code) {
                            case 1: --i; --j; state = 2; op = 0; break;
                            case 2: --i; state = 0; op = 1; break;
                            case 3: --i; state = 2; op = 1; break;
                            case 4: --j; state = 1; op = 2; break;
                            default: --j; state = 2; op = 2; break;
                        }
                        if (op == prevOp) ++run;
                        else {
                            if (nOps + 2 >= a.cigar_stage_cap) { status = PS_CIGAR_CAP; break; }
                            stage[nOps++] = cigar_pack((uint32_t)run, prevOp);
                            prevOp = op; run = 1;
                        }
                    }
                }
                i = __shfl_sync(FULL, i, 0);
                status = __shfl_sync(FULL, status, 0);
                __syncwarp();
            }
            if (lane == 0 && !status) {
                if (nOps + 2 >= a.cigar_stage_cap) status = PS_CIGAR_CAP;
                else if (op == 0) stage[nOps++] = cigar_pack((uint32_t)run + 1, 0);     // ssw.c:697-704
                else { stage[nOps++] = cigar_pack((uint32_t)run, op); stage[nOps++] = cigar_pack(1, 0); }
            }
            nOps = __shfl_sync(FULL, nOps, 0);
            status = __shfl_sync(FULL, status, 0);
        }
    }

    long long off = 0;
    if (!status) {
        if (lane == 0) off = (long long)atomicAdd(a.cigar_used, (unsigned long long)nOps);
        off = __shfl_sync(FULL, off, 0);
        __syncwarp();
        if (off + nOps <= a.cigar_cap) {
            for (int k = lane; k < nOps; k += 32) a.cigar_buf[off + k] = stage[nOps - 1 - k];
        } else status = PS_CIGAR_CAP;
    }
    if (lane == 0) {
        rec->cigar_off = off;
        rec->cigar_len = status ? 0 : nOps;
        rec->status |= status;
    }
}

template <int CLASS>
__global__ void __launch_bounds__(BAND_WARPS * 32, CLASS == 0 ? 4 : CLASS == 1 ? 2 : 1) band_kernel(const BandArgs a)
{
    __shared__ uint4 window[BAND_WARPS][TB_WINDOW / 16];
    __shared__ uint2 stab[8];
    const int count = *a.wl.count;
    if (count <= 0) return;
    if (threadIdx.x < 5) {
        // scores of read base rd = threadIdx.x against ref bases A C G T (bytes of .x) and N (.y)
        const int rd = threadIdx.x;
        unsigned lo = 0;
        for (int rf = 0; rf < 4; ++rf) lo |= (unsigned)(unsigned char)a.sc.mat[rf * 5 + rd] << (8 * rf);
        stab[rd] = make_uint2(lo, (unsigned)(unsigned char)a.sc.mat[4 * 5 + rd]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const int base = a.wl.base ? *a.wl.base : 0;
    unsigned char* ws = a.scratch + (size_t)(blockIdx.x * BAND_WARPS + warp) * a.scratch_stride;
    for (;;) {
        int idx = 0;
        if (lane_id() == 0) idx = atomicAdd(a.wl.cursor, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= count) break;
        band_pair<CLASS>(a, a.wl.idx[base + idx], ws, reinterpret_cast<unsigned char*>(window[warp]), stab);
        __syncwarp();
    }
}

cudaError_t launch_band(int cls, const BandArgs& a, int blocks, cudaStream_t st)
{
    if (cls == 0) band_kernel<0><<<blocks, BAND_WARPS * 32, 0, st>>>(a);
    else if (cls == 1) band_kernel<1><<<blocks, BAND_WARPS * 32, 0, st>>>(a);
    else band_kernel<2><<<blocks, BAND_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sswb
