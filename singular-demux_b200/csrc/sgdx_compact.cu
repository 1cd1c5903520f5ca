// sgdx_compact.cu -- format kernels of the compact records: one 64-bit word per read (barcodes of at most 21 bases).
//
//   bits  0..20  lo plane   bit j = bit 0 of base j's 2-bit code, code = (ascii >> 1) & 3  (A0 C1 T2 G3)
//   bits 21..41  hi plane   bit j = bit 1 of the code
//   bits 42..62  invalid    bit j = base j is not A/C/G/T; at such a position lo = 0 says 'N' (a no-call,
//                           demux.rs:78-80), lo = 1 says any other byte (mismatches every sample); hi = 0
//   bit  63      zero
//
// This is BASELINE.json's "2-bit base planes plus an N mask" and SURVEY 8(d)'s algorithmic input
// (ceil(3L/8) bytes per read): the host writes it instead of the L ASCII bytes when it fills the pinned batch
// buffer (sgdx_pack_compact_host, csrc/sgdx_hostpack.cpp).  Assignment on these records: sgdx_match_ph (sgdx_ph.cu)
// where the table allows its perfect hash, else the records are expanded to the 16-byte form of the general kernels.
#include "sgdx_seed_search.cuh"

namespace sgdx {

// ASCII -> compact records (device-resident pipelines and tests; the host packer writes the same words)
__global__ void __launch_bounds__(256) sgdx_pack_compact(const uint8_t *__restrict__ reads, uint64_t n, uint32_t L,
                                                         uint64_t *__restrict__ out) {
    const bool words_ok = (L & 3u) == 0u && (reinterpret_cast<uintptr_t>(reads) & 3u) == 0u;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x)
        out[r] = compact_of_rd(load_ascii(reads, r, L, words_ok));
}

// compact -> 16-byte sgdx_read records, for tables the fast kernel does not cover (no seed index, or
// min pairwise distance <= min_delta): those go through the packed form of the general kernels
__global__ void __launch_bounds__(256) sgdx_expand_compact(const uint64_t *__restrict__ in, uint64_t n, uint32_t lmask,
                                                           uint4 *__restrict__ out) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const Rd rd = rd_of_compact(__ldg(in + r), lmask);
        out[r] = make_uint4(rd.lo, rd.hi, rd.nmask, rd.xmask);
    }
}

cudaError_t launch_pack_compact(const uint8_t *reads, uint64_t n, uint32_t L, uint64_t *out, int sms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    sgdx_pack_compact<<<persistent_grid(n, 256, sms, 8), 256, 0, stream>>>(reads, n, L, out);
    return cudaGetLastError();
}

cudaError_t launch_expand_compact(const uint64_t *in, uint64_t n, uint32_t L, uint4 *out, int sms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    sgdx_expand_compact<<<persistent_grid(n, 256, sms, 8), 256, 0, stream>>>(in, n, (1u << L) - 1u, out);
    return cudaGetLastError();
}

}  // namespace sgdx
