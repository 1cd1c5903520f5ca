// sgdx_synth.cu -- seeded synthetic barcode generator, identical on host and device.
// The generator itself is sgdx_synth_core.h; this file adds the CUDA kernel and the C entry points.
#include "sgdx_guard.h"
#include "../../include/sgdx.h"
#include "../../include/sgdx_synth.h"
#include "sgdx_synth_core.h"

namespace sgdx {

__global__ void __launch_bounds__(256) synth_kernel(sgdx_synth_params p, const uint8_t *__restrict__ table,
                                                    uint64_t first, uint64_t n, uint8_t *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t tmp[kSynthMaxLen];
        synth_one(p, table, first + i, tmp);
        for (uint32_t j = 0; j < p.barcode_len; j++) out[i * p.barcode_len + j] = tmp[j];
    }
}

}  // namespace sgdx

using namespace sgdx;

extern "C" int sgdx_synth_table(uint64_t seed, uint32_t S, uint32_t L, uint32_t i1, uint32_t min_dist, uint8_t *out) {
    return sgdx::guarded("sgdx_synth_table", [&]() -> int {
    return synth_table_impl(seed, S, L, i1, min_dist, out) ? SGDX_EINVAL : SGDX_OK;
    });
}

extern "C" int sgdx_synth_reads_host(const sgdx_synth_params *p, const uint8_t *table, uint64_t first, uint64_t n,
                                     uint8_t *out, int threads) {
    return sgdx::guarded("sgdx_synth_reads_host", [&]() -> int {
    return synth_reads_host_impl(p, table, first, n, out, threads) ? SGDX_EINVAL : SGDX_OK;
    });
}

extern "C" int sgdx_synth_reads_device(const sgdx_synth_params *p, const uint8_t *d_table, uint64_t first, uint64_t n,
                                       uint8_t *d_out, void *stream) {
    return sgdx::guarded("sgdx_synth_reads_device", [&]() -> int {
    if (!p || !d_table || (n && !d_out) || p->barcode_len == 0 || p->barcode_len > SGDX_MAX_PACKED_LEN) return SGDX_EINVAL;
    if (n == 0) return SGDX_OK;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    synth_kernel<<<(uint32_t)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*p, d_table, first, n, d_out);
    return cudaGetLastError() == cudaSuccess ? SGDX_OK : SGDX_ECUDA;
    });
}
