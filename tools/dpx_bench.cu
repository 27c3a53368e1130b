// Microbenchmark: issue rates of the DPX s16x2 instructions and of the instruction mixes the SW
// inner loop can be built from (sm_100a).  Prints lane-instructions/s per GPU and per SM-clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_bench dpx_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int CH = 8;       // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(256) kern(unsigned* out, int iters, unsigned b, unsigned c, const unsigned* lut_g)
{
    __shared__ unsigned lut[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) lut[i] = lut_g[i];
    __syncthreads();
    unsigned a[CH], e[CH], q[CH];
    for (int k = 0; k < CH; ++k) { a[k] = threadIdx.x * 3 + k; e[k] = k * 7 + b; q[k] = ((threadIdx.x + k) & 63) * 128; }
    unsigned f = c, lane4 = (threadIdx.x & 31) * 4, base = 0;
    const char* lp = reinterpret_cast<const char*>(lut) + lane4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (MODE == 0) {            // pure VIADDMNMX, 5 per cell-pair, independent chains
                a[k] = __viaddmax_s16x2(a[k], b, c);
                e[k] = __viaddmax_s16x2(e[k], b, c);
                a[k] = __viaddmax_s16x2(a[k], c, b);
                e[k] = __viaddmax_s16x2(e[k], c, b);
                a[k] = __viaddmax_s16x2(a[k], b, e[k]);
            } else if (MODE == 1) {     // the Gotoh cell: 5 DPX, score from a register
                unsigned x = __viaddmax_s16x2(a[k], q[k], e[k]);
                unsigned h = __vimax_s16x2_relu(x, f);
                unsigned u = __viaddmax_s16x2(h, b, 0x80008000u);
                e[k] = __viaddmax_s16x2(e[k], c, u);
                f = __viaddmax_s16x2(f, c, u);
                a[k] = h;
            } else if (MODE == 2) {     // Gotoh cell + PRMT/XOR score (all ALU pipe)
                unsigned sel = q[k] ^ base;
                unsigned s = __byte_perm(b, c, sel);
                unsigned x = __viaddmax_s16x2(a[k], s, e[k]);
                unsigned h = __vimax_s16x2_relu(x, f);
                unsigned u = __viaddmax_s16x2(h, b, 0x80008000u);
                e[k] = __viaddmax_s16x2(e[k], c, u);
                f = __viaddmax_s16x2(f, c, u);
                a[k] = h;
            } else if (MODE == 3) {     // Gotoh cell + conflict-free LDS lookup (IADD + LDS)
                unsigned s = *reinterpret_cast<const unsigned*>(lp + ((q[k] + base) & 0x1f80));
                unsigned x = __viaddmax_s16x2(a[k], s, e[k]);
                unsigned h = __vimax_s16x2_relu(x, f);
                unsigned u = __viaddmax_s16x2(h, b, 0x80008000u);
                e[k] = __viaddmax_s16x2(e[k], c, u);
                f = __viaddmax_s16x2(f, c, u);
                a[k] = h;
            } else if (MODE == 4) {     // pure VIMNMX3
                a[k] = __vimax3_s16x2(a[k], b, e[k]);
                e[k] = __vimax3_s16x2(e[k], c, a[k]);
                a[k] = __vimax3_s16x2(a[k], c, b);
                e[k] = __vimax3_s16x2(e[k], b, c);
                a[k] = __vimax3_s16x2(a[k], e[k], c);
            } else if (MODE == 5) {     // 5 DPX + 2 IMAD (fma pipe) per pair: does the fma pipe run in the shadow?
                unsigned x = __viaddmax_s16x2(a[k], q[k], e[k]);
                unsigned h = __vimax_s16x2_relu(x, f);
                unsigned u = __viaddmax_s16x2(h, b, 0x80008000u);
                e[k] = __viaddmax_s16x2(e[k], c, u);
                f = __viaddmax_s16x2(f, c, u);
                q[k] = q[k] * b + c;
                a[k] = h * c + b;
            } else if (MODE == 6) {     // plain 32-bit IADD3/LOP3 ALU rate for reference
                a[k] = (a[k] + b) ^ c;
                e[k] = (e[k] + c) ^ b;
                a[k] = (a[k] + e[k]) ^ b;
                e[k] = (e[k] + b) ^ a[k];
                a[k] = (a[k] + c) ^ e[k];
            } else if (MODE == 7) {     // truncated-F cell: 6 DPX + LDS lookup
                unsigned s = *reinterpret_cast<const unsigned*>(lp + ((q[k] + base) & 0x1f80));
                unsigned x = __viaddmax_s16x2(a[k], s, e[k]);
                unsigned h0 = __viaddmax_s16x2_relu(f, q[k], x);
                unsigned h = __vimax_s16x2_relu(x, f);
                unsigned u = __viaddmax_s16x2(h0, b, 0x80008000u);
                e[k] = __viaddmax_s16x2(e[k], c, u);
                f = __viaddmax_s16x2(f, q[k], u);
                a[k] = h;
            }
        }
        base += 128;
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 7) f = __shfl_up_sync(0xffffffffu, f, 1);
    }
    unsigned r = f;
    for (int k = 0; k < CH; ++k) r ^= a[k] ^ e[k] ^ q[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
int run(const char* name, int dpx_per_unit, int other_per_unit, unsigned* out, const unsigned* lut, int sms, double clk_hz)
{
    const int iters = 4096, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<MODE><<<blocks, threads>>>(out, 64, 3, 5, lut);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        kern<MODE><<<blocks, threads>>>(out, iters, 0xffff0001u, 0xfffefffeu, lut);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double units = (double)blocks * threads * iters * CH;       // "cell-pair" units (one packed register)
    double dpx = units * dpx_per_unit / (best * 1e-3);
    double all = units * (dpx_per_unit + other_per_unit) / (best * 1e-3);
    printf("%-34s %8.3f ms  dpx lane-instr/s %.3e  (%.1f /clk/SM)  all lane-instr/s %.3e (%.1f /clk/SM)  packed-cells/s %.3e  GCUPS %.0f\n",
           name, best, dpx, dpx / clk_hz / sms, all, all / clk_hz / sms, units / (best * 1e-3), 2 * units / (best * 1e-3) / 1e9);
    return 0;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double clk = clk_khz * 1e3;
    printf("device %s sms %d clock %.0f MHz (per-clk figures assume max clock)\n", p.name, p.multiProcessorCount, clk / 1e6);
    unsigned* out; CK(cudaMalloc(&out, p.multiProcessorCount * 8 * 256 * 4));
    unsigned* lut; CK(cudaMalloc(&lut, 32 * 64 * 4)); CK(cudaMemset(lut, 1, 32 * 64 * 4));
    int sms = p.multiProcessorCount;
    run<0>("viaddmnmx x5 (independent)", 5, 0, out, lut, sms, clk);
    run<4>("vimnmx3 x5", 5, 0, out, lut, sms, clk);
    run<6>("iadd+lop x5 (10 alu)", 10, 0, out, lut, sms, clk);
    run<1>("gotoh cell 5 dpx", 5, 0, out, lut, sms, clk);
    run<2>("gotoh cell 5 dpx + xor + prmt", 5, 2, out, lut, sms, clk);
    run<3>("gotoh cell 5 dpx + iadd/and + lds", 5, 3, out, lut, sms, clk);
    run<5>("gotoh cell 5 dpx + 2 imad", 5, 2, out, lut, sms, clk);
    run<7>("trunc cell 6 dpx + iadd/and + lds", 6, 3, out, lut, sms, clk);
    return 0;
}
