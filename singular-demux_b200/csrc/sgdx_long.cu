// sgdx_long.cu -- sgdx_match_long: barcodes of 33..64 bases (sm_100a).
//
// The reference has no length cap (hamming_distance zips two slices of any length, matcher.rs:335-348); every other
// kernel of this library holds a barcode as 32-bit planes.  This one holds it as two halves of at most 32 bases --
// A = bases 0..31, B = bases 32..L-1 -- and is otherwise the per (read, sample) scan of sgdx_match_direct
// (sgdx_kernels.cu): ASCII input, one thread per read, the expected barcodes as {loA, hiA, loB, hiB} in shared memory
// (one 128-bit broadcast LDS per sample), distance = POPC over A + POPC over B with N / foreign bytes folded into the
// invalid masks, best / second-best keys, then the same rules (matcher.rs:266-303 / 121-262), the N gate and sample
// index (demux.rs:524-538), the per-sample counters (metrics.rs:428-437) and the index-hop check
// (metrics.rs:658-676; each index half is at most 32 bases, so the key sets of DevParams serve it).
// Long barcodes are rare (two 16+ base indexes plus an inline barcode); the kernel is the correct path for them, not a
// tuned one: 2 S POPC per read.
#include "sgdx_kernels.h"

namespace sgdx {

constexpr int kLongThreads = 256;

// `len` (<= 32) ASCII bytes at p -> planes.  words_ok: p 4-byte aligned and len % 4 == 0.
__device__ __forceinline__ Rd load_ascii_run(const uint8_t *__restrict__ p, uint32_t len, bool words_ok) {
    Rd rd = {0u, 0u, 0u, 0u};
    if (words_ok) {
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(p);
        const uint32_t nw = len >> 2;
#pragma unroll 8
        for (uint32_t i = 0; i < 8; i++)
            if (i < nw) fold_word(__ldg(pw + i), i * 4, rd);
    } else {
        for (uint32_t j = 0; j < len; j += 4) {
            uint32_t w = 0x41414141u;  // pad with 'A': contributes zero bits to every plane
            for (uint32_t k = 0; k < 4 && j + k < len; k++) w = (w & ~(0xFFu << (8 * k))) | ((uint32_t)__ldg(p + j + k) << (8 * k));
            fold_word(w, j, rd);
        }
    }
    const uint32_t inv = rd.nmask | rd.xmask;
    rd.lo &= ~inv;
    rd.hi &= ~inv;
    return rd;
}

// apply_rules (sgdx_device.cuh) with the no-call count and the foreign-byte flag of both halves given
template <uint32_t MODE>
__device__ __forceinline__ Verdict apply_rules_long(const DevParams &p, uint32_t n_no_calls, bool foreign, uint32_t best,
                                                    uint32_t nxt) {
    Verdict v;
    v.bin = p.S;
    v.dist = kNoDist;
    const bool gated = p.max_no_calls >= 0 && n_no_calls > (uint32_t)p.max_no_calls;  // demux.rs:528-531
    const uint32_t bd = best >> 16, nd = nxt >> 16;  // nd = 0xFFFF when there is no competitor
    bool ok;
    uint32_t reported;
    if (MODE == kModeHamming) {
        const uint32_t d = bd - min(n_no_calls, p.free_ns);  // matcher.rs:339-341
        reported = d;
        ok = best != kInfKey && d <= p.M && (nd - bd) > p.delta;  // matcher.rs:295-299
    } else {
        reported = bd;
        const bool rejected = nd <= bd + p.delta && nd <= p.pc_limit;  // visible competitor within delta
        ok = best != kInfKey && !foreign && bd <= p.M && !rejected;
    }
    if (ok && !gated) {
        v.bin = best & 0xFFFFu;
        v.dist = reported;
    }
    return v;
}

// planes of bases [from, from + len) of a barcode held as two halves, len <= 32
__device__ __forceinline__ uint32_t long_bits(uint32_t a, uint32_t b, uint32_t from) {
    return (uint32_t)((((uint64_t)b << 32) | a) >> from);
}

// metrics.rs:658-676 for one NoMatch read
__device__ __forceinline__ void hop_check_long(const DevParams &p, const Rd &A, const Rd &B) {
    const auto find = [&](const HopSlot *tab, uint32_t mask, uint32_t alln, uint32_t from, uint32_t len) {
        return hop_find(tab, mask, alln, long_bits(A.lo, B.lo, from), long_bits(A.hi, B.hi, from), long_bits(A.nmask, B.nmask, from),
                        long_bits(A.xmask, B.xmask, from), len);
    };
    const uint32_t a = find(p.hash1, p.hmask1, p.alln1, 0u, p.i1);
    if (a != 0xFFFFFFFFu) {
        const uint32_t b = find(p.hash2, p.hmask2, p.alln2, p.i1, p.i2);
        if (b != 0xFFFFFFFFu) {
            atomicAdd(&p.hop_fwd[(uint64_t)a * p.u2 + b], 1ull);
            return;
        }
    }
    // reversed orientation: index1 key in barcode[i2..], index2 key in barcode[..i2]
    const uint32_t c = find(p.hash1, p.hmask1, p.alln1, p.i2, p.i1);
    if (c == 0xFFFFFFFFu) return;
    const uint32_t d = find(p.hash2, p.hmask2, p.alln2, 0u, p.i2);
    if (d == 0xFFFFFFFFu) return;
    atomicAdd(&p.hop_rev[(uint64_t)c * p.u2 + d], 1ull);
}

// SMEM_TABLE: the expected barcodes fit shared memory next to the histogram (else they are read through L1 / L2)
template <uint32_t MODE, bool SMEM_TABLE>
__global__ void __launch_bounds__(kLongThreads) sgdx_match_long(const DevParams p, const uint4 *__restrict__ table) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw);  // (S+1)*3 counters, padded to 16 bytes
    const uint32_t nbins = (p.S + 1u) * 3u;
    uint4 *stab = reinterpret_cast<uint4 *>(smem_raw + (((size_t)nbins * 4u + 15u) & ~(size_t)15u));
    if (SMEM_TABLE)
        for (uint32_t i = threadIdx.x; i < p.S; i += blockDim.x) stab[i] = table[i];
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const uint4 *tab = SMEM_TABLE ? stab : table;

    const uint32_t LA = 32u, LB = p.L - 32u;
    const bool aligned = (reinterpret_cast<uintptr_t>(p.reads_ascii) & 3u) == 0u && (p.L & 3u) == 0u;  // then every run is
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = base + threadIdx.x;
        if (r >= p.n) continue;
        const uint8_t *row = p.reads_ascii + r * (uint64_t)p.L;
        const Rd A = load_ascii_run(row, LA, aligned), B = load_ascii_run(row + LA, LB, aligned);
        const uint32_t invA = A.nmask | A.xmask, invB = B.nmask | B.xmask;
        const uint32_t n_nc = __popc(A.nmask) + __popc(B.nmask), n_x = __popc(A.xmask) + __popc(B.xmask);
        // reads no table can match (cannot_match of sgdx_seed_search.cuh): every sample is further than M, or the gate
        bool dead = p.max_no_calls >= 0 && n_nc > (uint32_t)p.max_no_calls;
        if (MODE == kModeHamming) dead |= n_x + (n_nc > p.free_ns ? n_nc - p.free_ns : 0u) > p.M;
        else dead |= n_x != 0u || n_nc > p.M;
        // PreCompute: with a no-call in the read a competitor at exactly M+1 is not in the map (SURVEY A.3 iii)
        const uint32_t skip = (MODE == kModePreCompute && n_nc != 0u) ? p.M + 1u : 0xFFFFu;
        uint32_t best = kInfKey, nxt = kInfKey;
        if (!dead) {
#pragma unroll 2
            for (uint32_t s = 0; s < p.S; s++) {
                const uint4 t = tab[s];
                const uint32_t d = __popc(((A.lo ^ t.x) | invA) | (A.hi ^ t.y)) + __popc(((B.lo ^ t.z) | invB) | (B.hi ^ t.w));
                uint32_t key = (d << 16) | s;
                if (MODE == kModePreCompute) key = (d == skip) ? kInfKey : key;
                nxt = min(nxt, max(best, key));
                best = min(best, key);
            }
        }
        const Verdict v = apply_rules_long<MODE>(p, n_nc, n_x != 0u, best, nxt);
        p.out_idx[r] = (uint16_t)v.bin;
        if (p.out_dist) p.out_dist[r] = (uint8_t)v.dist;
        const uint32_t cls = v.dist == 0u ? 1u : (v.dist == 1u ? 2u : 0u);  // metrics.rs:428-437
        atomicAdd(&hist[v.bin * 3u + cls], 1u);
        if (p.hop_on && v.bin == p.S) hop_check_long(p, A, B);
    }

    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) {
        const uint32_t c = hist[i];
        if (c) atomicAdd(&p.counters[i], (unsigned long long)c);
    }
}

template <uint32_t MODE, bool SMEM_TABLE>
static cudaError_t launch_long_t(const DevParams &p, const uint4 *table, size_t smem, int sms, cudaStream_t stream) {
    auto kern = sgdx_match_long<MODE, SMEM_TABLE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLongThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint64_t tiles = (p.n + kLongThreads - 1) / kLongThreads, cap = (uint64_t)sms * (uint64_t)per_sm;
    kern<<<(uint32_t)(tiles < cap ? tiles : cap), kLongThreads, smem, stream>>>(p, table);
    return cudaGetLastError();
}

cudaError_t launch_long(const DevParams &p, const uint4 *table, uint32_t mode, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const size_t hist = (((size_t)(p.S + 1u) * 3u * 4u) + 15u) & ~(size_t)15u, with_table = hist + (size_t)p.S * 16u;
    const bool in_smem = with_table <= 100u * 1024u;  // two CTAs per SM; larger tables stay in L1 / L2
    const bool ham = mode == kModeHamming;
    if (in_smem)
        return ham ? launch_long_t<kModeHamming, true>(p, table, with_table, sms, stream)
                   : launch_long_t<kModePreCompute, true>(p, table, with_table, sms, stream);
    return ham ? launch_long_t<kModeHamming, false>(p, table, hist, sms, stream)
               : launch_long_t<kModePreCompute, false>(p, table, hist, sms, stream);
}

}  // namespace sgdx
