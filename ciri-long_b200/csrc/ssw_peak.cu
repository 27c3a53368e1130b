// ssw_peak.cu -- measures the issue rate of VIADDMNMX.S16x2 (the DPX instruction the score pass is
// built from) on the current device: the denominator of the DPX roofline.
// 8 independent dependency chains per thread, 8 CTAs of 256 threads per SM, best of 5 launches.
#include "ssw_common.cuh"
#include "ssw_kernels.h"

namespace sswb {

constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_OPS_PER_ITER = 5;

__global__ void __launch_bounds__(256) dpx_peak_kernel(unsigned* out, int iters, unsigned b, unsigned c)
{
    unsigned a[PEAK_CHAINS], e[PEAK_CHAINS];
    for (int k = 0; k < PEAK_CHAINS; ++k) { a[k] = threadIdx.x * 3 + k; e[k] = k * 7 + b; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PEAK_CHAINS; ++k) {
            a[k] = __viaddmax_s16x2(a[k], b, c);
            e[k] = __viaddmax_s16x2(e[k], b, c);
            a[k] = __viaddmax_s16x2(a[k], c, b);
            e[k] = __viaddmax_s16x2(e[k], c, b);
            a[k] = __viaddmax_s16x2(a[k], b, e[k]);
        }
    }
    unsigned r = 0;
    for (int k = 0; k < PEAK_CHAINS; ++k) r ^= a[k] ^ e[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

cudaError_t dpx_peak_probe(double* lane_instr_per_s, cudaStream_t st)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    unsigned* out = nullptr;
    e = cudaMalloc(&out, (size_t)blocks * threads * 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dpx_peak_kernel<<<blocks, threads, 0, st>>>(out, 64, 3, 5);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, st);
        dpx_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 0xffff0001u, 0xfffefffeu);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess) return e;
    const double ops = (double)blocks * threads * iters * PEAK_CHAINS * PEAK_OPS_PER_ITER;
    if (lane_instr_per_s) *lane_instr_per_s = ops / (best * 1e-3);
    return cudaGetLastError();
}

}  // namespace sswb
