// sgdx_single.cu -- assignment on compact records against ONE expected barcode (single-sample / inline-barcode runs,
// BASELINE config 4), no hop checker (sm_100a).
//
// Same results as every other assignment kernel (matcher.rs:273-348 / 121-262, demux.rs:524-538,
// metrics.rs:428-437).  With a single sample there is nothing to look up: the read's planes are compared with the
// barcode's (XOR / OR / POPC), the rule set is applied with "no competitor", and that is the verdict -- for every read,
// no-calls and foreign bytes included, so there is no slow path, no queue and no shared memory.  The kernel is a
// stream: one 256-bit load per four reads, one 64-bit and one 32-bit store for their results; the four counters
// (perfect, one mismatch, other match, undetermined) live in registers and reach the global counters once per warp.
#include "sgdx_seed_search.cuh"

namespace sgdx {

constexpr uint32_t kSingleThreads = 256;

template <uint32_t MODE>
__device__ __forceinline__ Verdict single_verdict(const DevParams &p, uint2 t, uint64_t rec, uint32_t lmask) {
    const Rd rd = rd_of_compact(rec, lmask);
    const uint32_t inv = rd.nmask | rd.xmask;
    const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
    uint32_t best = kInfKey, nxt = kInfKey;
    score<MODE>(rd, inv, skip, t.x, t.y, 0u, best, nxt);
    return apply_rules<MODE>(p, rd, best, nxt);
}

// VEC: n is a multiple of 4 and the three buffers are aligned for 256 / 64 / 32-bit accesses
template <uint32_t MODE, bool VEC>
__global__ void __launch_bounds__(kSingleThreads) sgdx_match_single(const DevParams p, const uint64_t *__restrict__ reads,
                                                                    uint32_t lmask) {
    const uint2 t = p.table[0];
    uint32_t cnt[4] = {0u, 0u, 0u, 0u};  // perfect, one mismatch, other match, undetermined
    const uint64_t stride = (uint64_t)gridDim.x * kSingleThreads;
    auto book = [&](const Verdict &v) {
        const uint32_t cls = v.bin ? 3u : min(v.dist, 2u);  // bin 0 = the sample, 1 = Undetermined
        cnt[0] += cls == 0u;
        cnt[1] += cls == 1u;
        cnt[2] += cls == 2u;
        cnt[3] += cls == 3u;
    };
    if (VEC) {
        const uint64_t groups = p.n / 4u;
        for (uint64_t g = (uint64_t)blockIdx.x * kSingleThreads + threadIdx.x; g < groups; g += stride) {
            uint64_t r[4];
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]) : "l"(reads + g * 4u));
            uint32_t bin[4], dist[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const Verdict v = single_verdict<MODE>(p, t, r[k], lmask);
                bin[k] = v.bin;
                dist[k] = v.dist;
                book(v);
            }
            *reinterpret_cast<uint2 *>(p.out_idx + g * 4u) = make_uint2(bin[0] | (bin[1] << 16), bin[2] | (bin[3] << 16));
            if (p.out_dist)
                *reinterpret_cast<uint32_t *>(p.out_dist + g * 4u) = dist[0] | (dist[1] << 8) | (dist[2] << 16) | (dist[3] << 24);
        }
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * kSingleThreads + threadIdx.x; i < p.n; i += stride) {
            const Verdict v = single_verdict<MODE>(p, t, __ldg(reads + i), lmask);
            p.out_idx[i] = (uint16_t)v.bin;
            if (p.out_dist) p.out_dist[i] = (uint8_t)v.dist;
            book(v);
        }
    }
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, cnt[c]);
        // counters are [bin][other, perfect, one mismatch]: the sample is bin 0, Undetermined bin 1 (class "other")
        const uint32_t at = c == 0u ? 1u : c == 1u ? 2u : c == 2u ? 0u : 3u;
        if ((threadIdx.x & 31u) == 0u && sum) atomicAdd(&p.counters[at], (unsigned long long)sum);
    }
}

template <uint32_t MODE, bool VEC>
static cudaError_t launch_single_t(const DevParams &p, const uint64_t *reads, uint32_t lmask, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const uint64_t units = VEC ? p.n / 4u : p.n;
    uint64_t grid = (units + kSingleThreads - 1u) / kSingleThreads;
    if (grid > (uint64_t)sms * 8u) grid = (uint64_t)sms * 8u;  // 8 CTAs of 256 threads per SM: full occupancy
    sgdx_match_single<MODE, VEC><<<(uint32_t)grid, kSingleThreads, 0, stream>>>(p, reads, lmask);
    return cudaGetLastError();
}

// aligned buffers: the multiple-of-4 part through the vector kernel, the last n % 4 reads through the scalar one
cudaError_t launch_single(const DevParams &p0, const uint64_t *reads, uint32_t mode, int sms, cudaStream_t stream) {
    DevParams p = p0;
    const uint32_t lmask = (1u << p.L) - 1u;
    const bool ham = mode == kModeHamming;
    const bool aligned = (reinterpret_cast<uintptr_t>(reads) & 31u) == 0u && (reinterpret_cast<uintptr_t>(p.out_idx) & 7u) == 0u &&
                         (reinterpret_cast<uintptr_t>(p.out_dist) & 3u) == 0u;
    if (aligned) {
        const uint64_t n_main = p.n & ~(uint64_t)3u;
        p.n = n_main;
        cudaError_t e = ham ? launch_single_t<kModeHamming, true>(p, reads, lmask, sms, stream)
                            : launch_single_t<kModePreCompute, true>(p, reads, lmask, sms, stream);
        if (e != cudaSuccess) return e;
        p.n = p0.n - n_main;
        reads += n_main;
        p.out_idx += n_main;
        if (p.out_dist) p.out_dist += n_main;
    }
    return ham ? launch_single_t<kModeHamming, false>(p, reads, lmask, sms, stream)
               : launch_single_t<kModePreCompute, false>(p, reads, lmask, sms, stream);
}

}  // namespace sgdx
