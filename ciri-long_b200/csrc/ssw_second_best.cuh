// ssw_second_best.cuh -- second-best scan shared by the packed and the 32-bit score kernels.
#pragma once
#include "ssw_common.cuh"

namespace sswb {

// Second-best score outside the mask window around the best end column (ssw.c:325-340 / 528-541).
// colbuf[c] = (colmax + off) | (H(last real row, c) + off) << 16.  P > 0: the reference's maxColumn[] also
// covers P zero-scoring pad rows that the DP did not compute.  A pad cell is reached from the last real
// row by free diagonal steps plus at most one horizontal gap, so
//   padmax(c) = max( max_{1<=d<=P} Hlast(c-d),  G(c) ),   G(c) = max(G(c-1) - ge, Hlast(c-P-1) - go, 0)
// (vertical gaps are dominated inside the same column).  Lanes take contiguous column chunks; the
// decaying chain G is stitched across chunks with one pass over the per-lane carries.
__device__ __noinline__ inline void second_best(const unsigned* colbuf, int n, int P, int off, int word, int endRef, int maskLen,
                                         int go, int ge, int lane, int& score2, int& ref2)
{
    const unsigned FULL = 0xffffffffu;
    const int e1 = endRef - maskLen > 0 ? endRef - maskLen : 0;              // left region  [0, e1)
    int e2 = endRef + maskLen > n ? n : endRef + maskLen;                    // right region [e2, n)
    if (!word) e2 += 1;                                                      // ssw.c:334 vs ssw.c:536
    const int chunk = (n + 31) / 32;
    const int c0 = lane * chunk < n ? lane * chunk : n;
    const int c1 = c0 + chunk < n ? c0 + chunk : n;

    int gin = 0;
    if (P > 0) {
        int carry = 0;
        for (int c = c0; c < c1; ++c) {
            const int src = c - P - 1 >= 0 ? (int)(colbuf[c - P - 1] >> 16) - off - go : -1;
            carry = carry - ge > src ? carry - ge : src;
            if (carry < 0) carry = 0;
        }
        int run = 0;
        for (int l = 0; l < 32; ++l) {
            const int cl = __shfl_sync(FULL, carry, l);
            const int len = __shfl_sync(FULL, c1 - c0, l);
            if (lane == l) gin = run;
            run = run - ge * len > cl ? run - ge * len : cl;
            if (run < 0) run = 0;
        }
    }
    int bv = 0, bi = 0x7fffffff, G = gin;
    for (int c = c0; c < c1; ++c) {
        int mc = (int)(colbuf[c] & 0xffffu) - off;
        if (P > 0) {
            const int src = c - P - 1 >= 0 ? (int)(colbuf[c - P - 1] >> 16) - off - go : -1;
            G = G - ge > src ? G - ge : src;
            if (G < 0) G = 0;
            int D = G;
            const int dmax = P < c ? P : c;
            for (int d = 1; d <= dmax; ++d) { const int v = (int)(colbuf[c - d] >> 16) - off; D = v > D ? v : D; }
            mc = D > mc ? D : mc;
        }
        if ((c < e1 || c >= e2) && mc > bv) { bv = mc; bi = c; }
    }
    const int M = __reduce_max_sync(FULL, bv);
    const int idx = __reduce_min_sync(FULL, (bv == M && M > 0) ? bi : 0x7fffffff);
    score2 = M;
    ref2 = M > 0 ? idx : 0;
}


}  // namespace sswb
