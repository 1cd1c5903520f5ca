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

// dense records (include/sgdx.h: 2 L bits per read, B = ceil(2 L / 8) bytes) -> compact records.  A thread takes four
// consecutive reads: their 4 B bytes are B aligned words; record k starts at bit 8 k B of that window.  The buffer is
// padded, so the words of the last (partial) group exist.
template <uint32_t B>
__global__ void __launch_bounds__(256) sgdx_unpack_dense(const uint32_t *__restrict__ dense, uint64_t n, uint32_t L,
                                                         uint64_t *__restrict__ out) {
    const uint64_t groups = (n + 3u) / 4u, fmask = (1ull << L) - 1ull;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w[B + 2];
#pragma unroll
        for (uint32_t i = 0; i < B; i++) w[i] = __ldg(dense + g * B + i);
        w[B] = w[B + 1] = 0u;
        uint64_t rec[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            constexpr uint32_t dummy = 0;
            (void)dummy;
            const uint32_t o = 8u * k * B, i = o >> 5, sh = o & 31u;  // compile-time after unrolling
            const uint32_t x0 = __funnelshift_r(w[i], w[i + 1], sh), x1 = __funnelshift_r(w[i + 1], w[i + 2], sh);
            const uint64_t x = (uint64_t)x0 | ((uint64_t)x1 << 32);
            rec[k] = (x & fmask) | (((x >> L) & fmask) << kCompactField);
        }
        if (4u * g + 3u < n) {
            reinterpret_cast<ulonglong2 *>(out + 4u * g)[0] = make_ulonglong2(rec[0], rec[1]);
            reinterpret_cast<ulonglong2 *>(out + 4u * g)[1] = make_ulonglong2(rec[2], rec[3]);
        } else {
            for (uint32_t k = 0; k < 4 && 4u * g + k < n; k++) out[4u * g + k] = rec[k];
        }
    }
}

// the same for barcodes of 22..32 bases: dense records -> 16-byte sgdx_read records {lo, hi, 0, 0}
template <uint32_t B>
__global__ void __launch_bounds__(256) sgdx_unpack_dense_wide(const uint32_t *__restrict__ dense, uint64_t n, uint32_t L,
                                                              uint4 *__restrict__ out) {
    const uint64_t groups = (n + 3u) / 4u, fmask = L >= 32u ? 0xFFFFFFFFull : (1ull << L) - 1ull;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w[B + 2];
#pragma unroll
        for (uint32_t i = 0; i < B; i++) w[i] = __ldg(dense + g * B + i);
        w[B] = w[B + 1] = 0u;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t o = 8u * k * B, i = o >> 5, sh = o & 31u;  // compile-time after unrolling
            const uint32_t x0 = __funnelshift_r(w[i], w[i + 1], sh), x1 = __funnelshift_r(w[i + 1], w[i + 2], sh);
            const uint64_t x = (uint64_t)x0 | ((uint64_t)x1 << 32);
            if (4u * g + k < n) out[4u * g + k] = make_uint4((uint32_t)(x & fmask), (uint32_t)((x >> L) & fmask), 0u, 0u);
        }
    }
}

__global__ void __launch_bounds__(256) sgdx_patch_exceptions_wide(const uint32_t *__restrict__ idx, const uint4 *__restrict__ rec,
                                                                  uint64_t n_exc, uint64_t n, uint4 *__restrict__ out) {
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_exc; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = __ldg(idx + e);
        if (r < n) out[r] = __ldg(rec + e);
    }
}

// the reads that are not plain A/C/G/T: their compact records over the placeholders
__global__ void __launch_bounds__(256) sgdx_patch_exceptions(const uint32_t *__restrict__ idx, const uint64_t *__restrict__ rec,
                                                             uint64_t n_exc, uint64_t n, uint64_t *__restrict__ out) {
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_exc; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = __ldg(idx + e);
        if (r < n) out[r] = __ldg(rec + e);
    }
}

cudaError_t launch_unpack_dense(const uint8_t *dense, uint64_t n, uint32_t L, const uint32_t *exc_idx, const uint64_t *exc_rec,
                                uint64_t n_exc, uint64_t *out, int sms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(dense);
    const uint32_t grid = persistent_grid((n + 3u) / 4u, 256, sms, 8);
    switch ((2u * L + 7u) / 8u) {
        case 1: sgdx_unpack_dense<1><<<grid, 256, 0, stream>>>(w, n, L, out); break;
        case 2: sgdx_unpack_dense<2><<<grid, 256, 0, stream>>>(w, n, L, out); break;
        case 3: sgdx_unpack_dense<3><<<grid, 256, 0, stream>>>(w, n, L, out); break;
        case 4: sgdx_unpack_dense<4><<<grid, 256, 0, stream>>>(w, n, L, out); break;
        case 5: sgdx_unpack_dense<5><<<grid, 256, 0, stream>>>(w, n, L, out); break;
        default: sgdx_unpack_dense<6><<<grid, 256, 0, stream>>>(w, n, L, out); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || n_exc == 0) return e;
    sgdx_patch_exceptions<<<persistent_grid(n_exc, 256, sms, 8), 256, 0, stream>>>(exc_idx, exc_rec, n_exc, n, out);
    return cudaGetLastError();
}

cudaError_t launch_unpack_dense_wide(const uint8_t *dense, uint64_t n, uint32_t L, const uint32_t *exc_idx, const uint4 *exc_rec,
                                     uint64_t n_exc, uint4 *out, int sms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(dense);
    const uint32_t grid = persistent_grid((n + 3u) / 4u, 256, sms, 8);
    switch ((2u * L + 7u) / 8u) {
        case 6: sgdx_unpack_dense_wide<6><<<grid, 256, 0, stream>>>(w, n, L, out); break;
        case 7: sgdx_unpack_dense_wide<7><<<grid, 256, 0, stream>>>(w, n, L, out); break;
        default: sgdx_unpack_dense_wide<8><<<grid, 256, 0, stream>>>(w, n, L, out); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || n_exc == 0) return e;
    sgdx_patch_exceptions_wide<<<persistent_grid(n_exc, 256, sms, 8), 256, 0, stream>>>(exc_idx, exc_rec, n_exc, n, out);
    return cudaGetLastError();
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
