// tools/tband_host_check.cpp -- TEST INFRASTRUCTURE.  Runs the lane program of the throughput CIGAR kernel
// (ciri-long_b200/csrc/ssw_tband_core.h) on the host, 32 emulated lanes in lock step exactly like the device
// driver loop of ssw_tband.cu, and compares every CIGAR with oracle/ssw_oracle.c (orc_align).
//   g++ -O2 -o /tmp/tband_check tools/tband_host_check.cpp oracle/ssw_oracle.c -I. && /tmp/tband_check [pairs] [seed]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../ciri-long_b200/csrc/ssw_tband_core.h"
extern "C" {
#include "../oracle/ssw_oracle.h"
}
using namespace sswt;

struct Pair { std::vector<int8_t> q, r; orc_result res; int st; };

static unsigned rng_state = 12345;
static long g_blocks_masked = 0, g_blocks_plain = 0, g_quirk_rows = 0, g_passes = 0, g_mode[4] = {0, 0, 0, 0};
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

struct LaneOut { int nOps; std::vector<unsigned> ops; int bwFinal; };

// one warp of up to 32 jobs, all passes (the device re-sorts between passes; here a lane simply keeps its pair)
static void run_warp(std::vector<TbJob>& jobs, const int8_t* mat, int go, int ge, std::vector<LaneOut>& out)
{
    const int n = (int)jobs.size();
    const int B = go + ge;
    const unsigned B2 = (unsigned)B | ((unsigned)B << 16), GO2 = (unsigned)go * 0x10001u, GE2 = (unsigned)ge * 0x10001u;
    int matS[25]; for (int k = 0; k < 25; ++k) matS[k] = mat[k];
    out.assign(n, LaneOut());
    std::vector<char> done(n, 0);
    for (;;) {
        bool any = false; for (int l = 0; l < n; ++l) any |= !done[l];
        if (!any) break;
        ++g_passes;
        int NB = 1, rowPairsMax = 0;
        for (int l = 0; l < n; ++l) if (!done[l]) { NB = std::max(NB, tb_blocks(jobs[l].bw)); rowPairsMax = std::max(rowPairsMax, (jobs[l].readLen + 1) / 2); }
        const int RING = 4 * NB;
        std::vector<unsigned> S((size_t)(4 * NB + 4) * 32, B2);
        std::vector<unsigned char> ring((size_t)(RING + 4) * 32, 0);
        std::vector<int> tab(10 * 32, 0);
        std::vector<unsigned> dirs((size_t)rowPairsMax * NB * 32, 0);
        std::vector<unsigned> maxv2(n), finalMax(n, 0);
        auto refcode = [&](const TbJob& J, int col) { return (unsigned char)tb_ref_code(J, col); };
        for (int l = 0; l < n; ++l) {
            if (done[l]) continue;
            const TbJob& J = jobs[l];
            for (int p = 0; p < RING; ++p) ring[(size_t)p * 32 + l] = refcode(J, p - J.bw);
            for (int p = 0; p < 4; ++p) ring[(size_t)(RING + p) * 32 + l] = ring[(size_t)p * 32 + l];
            maxv2[l] = (unsigned)(J.maxIn + B) * 0x10001u;
        }
        std::vector<TbRow> R(n);
        for (int rho = 0; rho < rowPairsMax; ++rho) {
            int head = 0, tail = 4 * NB; bool simple = true;
            for (int l = 0; l < n; ++l) {
                if (done[l]) continue;
                tb_row_begin(R[l], jobs[l], rho, &S[l], &tab[l], &tab[l], matS, B2, tb_read_code(jobs[l], 2 * rho), tb_read_code(jobs[l], 2 * rho + 1));
                if (rho < (jobs[l].readLen + 1) / 2) {
                    const TbRowPlan pl = tb_row_plan(R[l], NB);
                    if (R[l].tqL >= 0 || R[l].tqH >= 0) ++g_quirk_rows;
                    head = std::max(head, pl.lo); tail = std::min(tail, pl.hi + 1); simple = simple && pl.simple;
                }
            }
            int hb = (head + 3) / 4, tb = tail / 4;
            if (tb < hb) { hb = NB; tb = NB; }
            const bool headMode = simple && hb == 1;
            const int pos0row = (2 * rho) % RING;
            for (int b = 0; b < NB; ++b) {
                const int pos0 = (pos0row + 4 * b) % RING;
                const int mode = b < hb ? (headMode ? TB_HEAD : TB_ANY) : b < tb ? TB_PLAIN : (hb == NB ? TB_ANY : TB_TAIL);
                (mode != TB_PLAIN ? g_blocks_masked : g_blocks_plain) += 1; g_mode[mode] += 1;
                for (int l = 0; l < n; ++l) {
                    if (done[l]) continue;
                    unsigned w;
                    const unsigned char* rp = &ring[(size_t)pos0 * 32 + l];
                    if (mode == TB_ANY) w = tb_block<TB_ANY>(R[l], 4 * b, &S[l], rp, &tab[l], B2, GO2, GE2, 1u, maxv2[l]);
                    else if (mode == TB_HEAD) w = tb_block<TB_HEAD>(R[l], 4 * b, &S[l], rp, &tab[l], B2, GO2, GE2, 1u, maxv2[l]);
                    else if (mode == TB_TAIL) w = tb_block<TB_TAIL>(R[l], 4 * b, &S[l], rp, &tab[l], B2, GO2, GE2, 1u, maxv2[l]);
                    else w = tb_block<TB_PLAIN>(R[l], 4 * b, &S[l], rp, &tab[l], B2, GO2, GE2, 1u, maxv2[l]);
                    dirs[((size_t)rho * NB + b) * 32 + l] = w;
                }
            }
            for (int l = 0; l < n; ++l) {
                if (done[l]) continue;
                const TbJob& J = jobs[l];
                for (int k = 0; k < 2; ++k) {
                    const int pos = (pos0row + k) % RING;
                    const unsigned char c = refcode(J, 2 * rho + RING + k - J.bw);
                    ring[(size_t)pos * 32 + l] = c;
                    if (pos < 4) ring[(size_t)(RING + pos) * 32 + l] = c;
                }
                if (rho + 1 == (J.readLen + 1) / 2) finalMax[l] = maxv2[l];
            }
        }
        for (int l = 0; l < n; ++l) {
            if (done[l]) continue;
            TbJob& J = jobs[l];
            const short lo = (short)(finalMax[l] & 0xffff), hi = (short)(finalMax[l] >> 16);
            const int mx = (lo > hi ? lo : hi) - B;
            if (mx < J.score && 2 * J.bw < 2 * J.readLen) { J.bw *= 2; J.maxIn = mx; if (tb_steps(J.bw) > TB_MAX_STEPS) { done[l] = 1; out[l].nOps = -9; } continue; }
            std::vector<unsigned> stage((size_t)(J.readLen + J.refLen + 8) * 32, 0);
            const int nOps = tb_traceback(J, &dirs[l], NB, &stage[l], J.readLen + J.refLen + 8);
            out[l].nOps = nOps; out[l].bwFinal = J.bw;
            for (int k = 0; k < nOps; ++k) out[l].ops.push_back(stage[(size_t)(nOps - 1 - k) * 32 + l]);
            done[l] = 1;
        }
    }
}

int main(int argc, char** argv)
{
    const int total = argc > 1 ? atoi(argv[1]) : 2000;
    rng_state = argc > 2 ? (unsigned)atoi(argv[2]) : 12345u;
    const int schemes[4][4] = {{1, 1, 1, 1}, {10, 4, 8, 2}, {2, 2, 3, 1}, {2, 2, 2, 2}};
    long checked = 0, bad = 0, skipped = 0, tberr = 0;
    for (int done = 0; done < total;) {
        const int* sc = schemes[rnd() % 4];
        int8_t mat[25];
        for (int a = 0; a < 5; ++a) for (int b = 0; b < 5; ++b) mat[a * 5 + b] = (a == 4 || b == 4) ? 0 : (a == b ? sc[0] : -sc[1]);
        const int n = 1 + rnd() % 32;
        const int family = rnd() % 5;        // 0 tiny, 1 medium, 2 C2-like, 3 short-vs-long (wide bands), 4 one length per warp (sorted-list case)
        const int m4 = 40 + rnd() % 400;
        std::vector<Pair> pairs(n);
        for (auto& p : pairs) {
            int m = family == 4 ? m4 : family == 0 ? 8 + rnd() % 50 : family == 1 ? 60 + rnd() % 200 : family == 2 ? 250 + rnd() % 300 : 20 + rnd() % 120;
            std::vector<int8_t> core(m);
            for (auto& c : core) c = rnd() % 4;
            p.q.clear(); p.r.clear();
            const int fl = family == 4 ? 5 : rnd() % 30, fr = family == 4 ? 5 : rnd() % 30;
            for (int k = 0; k < fl; ++k) p.r.push_back(rnd() % 4);
            const int err = family == 3 ? 25 : (rnd() % 3 == 0 ? 20 : 8);
            for (int k = 0; k < m; ++k) {
                const unsigned u = rnd() % 100;
                if ((int)u < err / 3) { p.q.push_back(core[k]); }                                    // deletion in ref
                else if ((int)u < 2 * err / 3) { p.q.push_back(core[k]); p.r.push_back(core[k]); const int run = 1 + rnd() % (family == 3 ? 12 : 3); for (int z = 0; z < run; ++z) p.r.push_back(rnd() % 4); }
                else if ((int)u < err) { p.q.push_back(core[k]); p.r.push_back((core[k] + 1 + rnd() % 3) % 4); }
                else { p.q.push_back(core[k]); p.r.push_back(core[k]); }
                if (rnd() % 100 == 0) p.q.back() = 4;
            }
            for (int k = 0; k < fr; ++k) p.r.push_back(rnd() % 4);
            if (rnd() % 2) std::swap(p.q, p.r);
            const int ml = (int)p.q.size() > 30 ? (int)p.q.size() / 2 : 15;
            p.st = orc_align(p.q.data(), (int)p.q.size(), p.r.data(), (int)p.r.size(), mat, 5, 2, sc[2], sc[3], 1, 0, 0, ml, &p.res);
        }
        std::vector<TbJob> jobs; std::vector<int> who;
        for (int k = 0; k < n; ++k) {
            Pair& p = pairs[k];
            if (p.res.ref_begin1 < 0 || p.res.score1 + sc[2] + sc[3] >= TB_SCORE_LIMIT) { ++skipped; continue; }
            TbJob J;
            J.ref = p.r.data() + p.res.ref_begin1; J.read = p.q.data() + p.res.read_begin1;
            J.refLen = p.res.ref_end1 - p.res.ref_begin1 + 1; J.readLen = p.res.read_end1 - p.res.read_begin1 + 1;
            J.bw = abs(J.refLen - J.readLen) + 1; J.score = p.res.score1; J.maxIn = 0;
            if (tb_steps(J.bw) > TB_MAX_STEPS) { ++skipped; continue; }
            jobs.push_back(J); who.push_back(k);
        }
        std::vector<LaneOut> out;
        if (!jobs.empty()) run_warp(jobs, mat, sc[2], sc[3], out);
        for (size_t x = 0; x < jobs.size(); ++x) {
            Pair& p = pairs[who[x]];
            if (out[x].nOps == -9) { ++skipped; continue; }
            ++checked;
            bool ok;
            if (p.st == ORC_ERR_TRACEBACK) { ok = out[x].nOps == -1; ++tberr; }
            else {
                ok = p.st == ORC_OK && out[x].nOps == p.res.cigarLen && out[x].bwFinal == p.res.band_width;
                for (int k = 0; ok && k < p.res.cigarLen; ++k) ok = out[x].ops[k] == p.res.cigar[k];
            }
            if (!ok) {
                if (++bad <= 5) {
                    fprintf(stderr, "MISMATCH scheme %d/%d/%d/%d refLen %d readLen %d score %d: got nOps %d bw %d, want %d bw %d st %d\n",
                            sc[0], sc[1], sc[2], sc[3], jobs[x].refLen, jobs[x].readLen, p.res.score1, out[x].nOps, out[x].bwFinal, p.res.cigarLen, p.res.band_width, p.st);
                    for (int k = 0; k < out[x].nOps && k < 40; ++k) fprintf(stderr, "%u%c", out[x].ops[k] >> 4, "MID"[out[x].ops[k] & 3]); fprintf(stderr, "\n");
                    for (int k = 0; k < p.res.cigarLen && k < 40; ++k) fprintf(stderr, "%u%c", p.res.cigar[k] >> 4, "MID"[p.res.cigar[k] & 3]); fprintf(stderr, "\n");
                }
            }
        }
        for (auto& p : pairs) orc_result_free(&p.res);
        done += n;
    }
    printf("tband host check: %ld pairs checked, %ld mismatches, %ld skipped, %ld traceback-error cases\n", checked, bad, skipped, tberr);
    printf("  warp passes %ld, blocks masked %ld plain %ld, quirk row pairs %ld; modes plain/any/head/tail %ld/%ld/%ld/%ld\n", g_passes, g_blocks_masked, g_blocks_plain, g_quirk_rows, g_mode[0], g_mode[1], g_mode[2], g_mode[3]);
    return bad ? 1 : 0;
}
