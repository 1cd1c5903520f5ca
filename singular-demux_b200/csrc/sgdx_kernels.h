// sgdx_kernels.h -- launcher interface between the C-ABI layer and the kernels.
#pragma once
#include "sgdx_device.cuh"

namespace sgdx {

constexpr int kDirectThreads = 256;
constexpr int kSeededThreads = 256;

size_t direct_smem_bytes(uint32_t S);
cudaError_t launch_direct(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream);
size_t seeded_smem_bytes(const DevParams &p);
size_t exact_smem_bytes(const DevParams &p, uint32_t threads);
uint32_t exact_block_threads(const DevParams &p);
cudaError_t launch_seeded(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream);
cudaError_t launch_pack(const uint8_t *reads, uint64_t n, uint32_t L, uint4 *out, int sms, cudaStream_t stream);

// integer-pipe micro-benchmark (sgdx_microbench.cu)
cudaError_t measure_int_peak(int sms, double *popc_per_s, double *lop3_per_s);

}  // namespace sgdx
