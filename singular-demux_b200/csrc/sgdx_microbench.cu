// sgdx_microbench.cu -- measures the integer-pipe peaks the assignment kernels are bounded by:
// sustained 32-bit POPC results/s and LOP3 results/s over the whole chip (MEASURED_PEAKS.json has no
// integer figure, SURVEY 8(d) asks for one measured on the box).
#include "sgdx_kernels.h"

namespace sgdx {

constexpr int kChains = 8;
constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) popc_chain(uint32_t *out, uint32_t seed) {
    uint32_t a[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) a[c] = seed * (threadIdx.x + 1u) + blockIdx.x + c * 0x9E3779B9u;
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int c = 0; c < kChains; c++) {
            uint32_t t;
            asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(a[c]));
            a[c] = t | seed;  // LOP on the (much wider) alu pipe keeps the chain data-dependent
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) s ^= a[c];
    if (s == 0x12345678u) out[0] = s;
}

__global__ void __launch_bounds__(256) lop3_chain(uint32_t *out, uint32_t seed) {
    uint32_t a[kChains];
    uint32_t b = seed ^ threadIdx.x, d = seed * 3u + blockIdx.x;
#pragma unroll
    for (int c = 0; c < kChains; c++) a[c] = seed * (threadIdx.x + 1u) + c * 0x9E3779B9u;
    for (int i = 0; i < kIters; i++) {
#pragma unroll
        for (int c = 0; c < kChains; c++)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b), "r"(d));  // a ^ b ^ d
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) s ^= a[c];
    if (s == 0x12345678u) out[0] = s;
}

template <typename K>
static cudaError_t time_kernel(K kern, int grid, uint32_t *d_out, double *ops_per_s) {
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        kern<<<grid, 256>>>(d_out, 12345u + rep);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) return e;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_s = (double)grid * 256.0 * kChains * kIters / (best * 1e-3);
    return cudaGetLastError();
}

cudaError_t measure_int_peak(int sms, double *popc_per_s, double *lop3_per_s) {
    uint32_t *d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_out, 4);
    if (e != cudaSuccess) return e;
    int grid = sms * 8;
    e = time_kernel(popc_chain, grid, d_out, popc_per_s);
    if (e == cudaSuccess) e = time_kernel(lop3_chain, grid, d_out, lop3_per_s);
    cudaFree(d_out);
    return e;
}

}  // namespace sgdx
