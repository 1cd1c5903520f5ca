// sgdx_kernels.cu -- hand-written sm_100a kernels of the sample-barcode assignment path.
//
//   sgdx_match_direct  : the per (read, sample) XOR/OR/POPC scan with best/second-best keys.
//                        Replaces find_independently + hamming_distance (matcher.rs:273-348), the
//                        PreCompute lookup (matcher.rs:129-140 via the closed form of its map), the N gate
//                        and index selection (demux.rs:524-538), update_with_match (metrics.rs:428-437)
//                        and check_barcode (metrics.rs:658-676).
//   sgdx_pack          : ASCII -> 16-byte plane records (128-bit stores).
#include "sgdx_kernels.h"

namespace sgdx {

// ------------------------------------------------------------------------------------------------
// direct kernel: one thread per read, expected-barcode table in shared memory (two samples per
// 128-bit broadcast LDS), N-always-mismatch distance so the inner loop is 2 LOP3 + 1 POPC per pair.
// ------------------------------------------------------------------------------------------------
template <uint32_t MODE, bool PACKED>
__global__ void __launch_bounds__(kDirectThreads) sgdx_match_direct(const DevParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *stab = reinterpret_cast<uint4 *>(smem_raw);                      // table_pad / 2 pairs
    uint32_t *hist = reinterpret_cast<uint32_t *>(stab + p.table_pad / 2);  // (S+1)*3 counters
    const uint32_t nbins = (p.S + 1u) * 3u;

    for (uint32_t i = threadIdx.x; i < p.table_pad / 2; i += blockDim.x) {
        uint2 a = p.table[2 * i];
        uint2 b = (2 * i + 1 < p.S) ? p.table[2 * i + 1] : make_uint2(0u, 0u);
        stab[i] = make_uint4(a.x, a.y, b.x, b.y);
    }
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0u;
    __syncthreads();

    const bool words_ok = (p.L & 3u) == 0u && (reinterpret_cast<uintptr_t>(p.reads_ascii) & 3u) == 0u;
    const uint32_t npairs = p.S >> 1;

    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = base + threadIdx.x;
        const bool active = r < p.n;
        Rd rd = {0u, 0u, 0u, 0u};
        if (active) rd = PACKED ? load_packed(p.reads_packed, r, p.L) : load_ascii(p.reads_ascii, r, p.L, words_ok);
        const uint32_t inv = rd.nmask | rd.xmask;
        // PreCompute: with a no-call in the read a competitor at exactly M+1 is not in the map (A.3 iii)
        const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;

        uint32_t best = kInfKey, nxt = kInfKey;
        auto consider = [&](uint32_t slo, uint32_t shi, uint32_t s) {
            // two 3-input LOP3s: t = (lo ^ slo) | inv ; m = t | (hi ^ shi)
            uint32_t t, m;
            asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(t) : "r"(rd.lo), "r"(slo), "r"(inv));
            asm("lop3.b32 %0, %1, %2, %3, 0xF6;" : "=r"(m) : "r"(t), "r"(rd.hi), "r"(shi));
            uint32_t d = __popc(m);
            uint32_t key;  // (d << 16) + s as one IMAD on the fma pipe (the alu pipe is the busy one)
            asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(key) : "r"(d), "r"(s));
            if (MODE == kModePreCompute) key = (d == skip) ? kInfKey : key;
            nxt = min(nxt, max(best, key));
            best = min(best, key);
        };
#pragma unroll 4
        for (uint32_t i = 0; i < npairs; i++) {
            uint4 t = stab[i];
            consider(t.x, t.y, 2 * i);
            consider(t.z, t.w, 2 * i + 1);
        }
        if (p.S & 1u) {
            uint4 t = stab[npairs];
            consider(t.x, t.y, p.S - 1u);
        }
        Verdict v = apply_rules<MODE>(p, rd, best, nxt);
        emit(p, r, active, rd, v, hist);
    }

    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) {
        uint32_t c = hist[i];
        if (c) atomicAdd(&p.counters[i], (unsigned long long)c);
    }
}

// ------------------------------------------------------------------------------------------------
// pack kernel: ASCII -> sgdx_read records
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgdx_pack(const uint8_t *__restrict__ reads, uint64_t n, uint32_t L,
                                                 uint4 *__restrict__ out) {
    const bool words_ok = (L & 3u) == 0u && (reinterpret_cast<uintptr_t>(reads) & 3u) == 0u;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        Rd rd = load_ascii(reads, r, L, words_ok);
        out[r] = make_uint4(rd.lo, rd.hi, rd.nmask, rd.xmask);
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static uint32_t grid_for(uint64_t n, uint32_t threads, int sms, int blocks_per_sm) {
    uint64_t tiles = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)sms * (uint64_t)blocks_per_sm;  // persistent grid: a multiple of the SM count
    return (uint32_t)(tiles < cap ? (tiles ? tiles : 1) : cap);
}

size_t direct_smem_bytes(uint32_t S) {
    uint32_t pad = (S + 1u) & ~1u;
    return (size_t)pad * 8u + (size_t)(S + 1u) * 3u * 4u;
}

template <uint32_t MODE, bool PACKED>
static cudaError_t launch_direct_t(const DevParams &p, int sms, cudaStream_t stream) {
    size_t smem = direct_smem_bytes(p.S);
    auto kern = sgdx_match_direct<MODE, PACKED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDirectThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = grid_for(p.n, kDirectThreads, sms, per_sm);
    kern<<<grid, kDirectThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_direct(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    if (mode == kModeHamming)
        return packed ? launch_direct_t<kModeHamming, true>(p, sms, stream)
                      : launch_direct_t<kModeHamming, false>(p, sms, stream);
    return packed ? launch_direct_t<kModePreCompute, true>(p, sms, stream)
                  : launch_direct_t<kModePreCompute, false>(p, sms, stream);
}

cudaError_t launch_pack(const uint8_t *reads, uint64_t n, uint32_t L, uint4 *out, int sms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint32_t grid = grid_for(n, 256, sms, 8);
    sgdx_pack<<<grid, 256, 0, stream>>>(reads, n, L, out);
    return cudaGetLastError();
}

}  // namespace sgdx
