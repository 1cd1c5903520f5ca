// sgdx_ph.cu -- assignment on compact records through a perfect hash of the match neighbourhood (sm_100a).
//
// Same results as every other assignment kernel (matcher.rs:273-348 / 121-262, demux.rs:524-538,
// metrics.rs:428-437, 658-676); input = one 64-bit compact record per read (include/sgdx.h).
//
// Idea.  For a read made of A/C/G/T only, Matcher::find is a pure function of the string.  Only strings within
// max_mismatches substitutions of some expected barcode can be a Match, and there are few of them
// (S * sum_k C(L,k) 3^k: 23 424 for 384 samples of 20 bases at one mismatch).  sgdx_create enumerates that key set,
// decides each key with a full scan under the handle's rule set (best / second best, ties, delta, PreCompute's
// visibility rules -- exactly what the general kernels do per read), and stores "sample or none" in a two-level
// perfect hash (PhParams, sgdx_kernels.h).  Per read the kernel then does
//     three multiplies -> bucket multiplier (LDS.U16) -> entry (LDS.U16) -> that sample's planes (LDS.64)
//     -> XOR / OR / POPC -> distance <= max_mismatches ?
// The distance test is both the verification of the untagged entry and the reported hamming_dist.  There is no
// search, no queue and no barrier on this path; every read that is not "A/C/G/T only" -- and the index-hop check of
// the NoMatch reads -- is deferred to per-warp queues in shared memory and drained 32 at a time by the full warp.
//
// Layout.  One 1024-thread CTA per SM, warps are independent (no __syncthreads between the table load and the final
// counter flush).  A warp takes chunks of 128 consecutive reads: lane i owns reads 4i..4i+3 of the chunk, fetched by
// one 256-bit load, answered by one 64-bit store (4 sample indices) and one 32-bit store (4 distances).  The next
// chunk's records are requested before the current chunk is processed.  Buffers that are not aligned for this, and
// the last partial chunk, take the same code with 64-bit loads and 16/8-bit stores (lane i owns reads i, i+32, ...).
#include "sgdx_seed_search.cuh"

namespace sgdx {

constexpr uint32_t kPhThreads = 1024;
constexpr uint32_t kPhWarps = kPhThreads / 32;
constexpr uint32_t kPhChunk = 128;  // reads per warp per iteration
constexpr uint32_t kPhQueue = 160;  // <= 31 left over + <= 128 pushed by one chunk

struct PhSmem {
    uint16_t *disp;   // [1 << bbits]
    uint16_t *slots;  // [1 << sbits]
    uint2 *planes;    // [S + 1], entry S = dummy that nothing matches
    uint32_t *hist;   // [3][S + 1]: perfect, one mismatch, other (metrics.rs:428-437)
    uint64_t *qhop;   // [warps][kPhQueue] records of A/C/G/T-only NoMatch reads awaiting the hop check
    uint32_t *qslow;  // [warps][kPhQueue] (warp iteration * 128 + offset in chunk) of reads for the general path
};

__host__ __device__ inline size_t ph_carve(unsigned char *base, uint32_t S, uint32_t bbits, uint32_t sbits, PhSmem *out) {
    size_t o = 0;
    PhSmem s;
    s.qhop = reinterpret_cast<uint64_t *>(base + o);   o += (size_t)kPhWarps * kPhQueue * 8;
    s.planes = reinterpret_cast<uint2 *>(base + o);    o += (((size_t)S + 1) * 8 + 15) & ~(size_t)15;
    s.hist = reinterpret_cast<uint32_t *>(base + o);   o += (((size_t)S + 1) * 3 * 4 + 15) & ~(size_t)15;
    s.qslow = reinterpret_cast<uint32_t *>(base + o);  o += (size_t)kPhWarps * kPhQueue * 4;
    s.disp = reinterpret_cast<uint16_t *>(base + o);   o += (size_t)2 << bbits;
    s.slots = reinterpret_cast<uint16_t *>(base + o);  o += (size_t)2 << sbits;
    o = (o + 15) & ~(size_t)15;
    if (out) *out = s;
    return o;
}

size_t ph_smem_bytes(uint32_t S, uint32_t bbits, uint32_t sbits) { return ph_carve(nullptr, S, bbits, sbits, nullptr); }

// the perfect-hash probe of one record without invalid positions: entry (sample or S) and distance to it
__device__ __forceinline__ void ph_probe(const PhSmem &sm, const PhParams &c, uint32_t w0, uint32_t w1, uint32_t &e,
                                         uint32_t &d) {
    const uint32_t h = __umulhi(w0, c.k0) + w0 * c.k1 + w1 * c.k2;
    const uint32_t mul = sm.disp[h >> c.bshift];
    e = sm.slots[(h * mul) >> c.sshift];
    const uint2 t = sm.planes[e];
    const uint32_t lo = w0 & kCompactMask;
    const uint32_t hi = __funnelshift_r(w0, w1, kCompactField);  // bits 21..52 of the record; 42..52 are zero here
    d = (uint32_t)__popc((lo ^ t.x) | (hi ^ t.y));
}

// verdict of one read of the general path: patch the placeholder the fast path stored, count, hop check
__device__ __forceinline__ void ph_emit(const DevParams &p, uint32_t *hist, uint32_t S1, uint64_t r, const Rd &rd,
                                        const Verdict &v) {
    p.out_idx[r] = (uint16_t)v.bin;
    if (p.out_dist) p.out_dist[r] = (uint8_t)v.dist;
    atomicAdd(&hist[min(v.dist, 2u) * S1 + v.bin], 1u);  // metrics.rs:428-437
    if (p.hop_on && v.bin == p.S) hop_check(p, rd);
}

template <uint32_t MODE>
__global__ void __launch_bounds__(kPhThreads, 1) sgdx_match_ph(const DevParams p, const PhParams c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhSmem sm;
    const uint32_t bbits = 32u - c.bshift, sbits = 32u - c.sshift;
    ph_carve(smem_raw, p.S, bbits, sbits, &sm);
    const uint32_t S1 = p.S + 1u, nbins = S1 * 3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

    for (uint32_t i = threadIdx.x; i < (1u << bbits); i += kPhThreads) sm.disp[i] = c.disp[i];
    for (uint32_t i = threadIdx.x; i < (1u << sbits); i += kPhThreads) sm.slots[i] = c.slots[i];
    for (uint32_t i = threadIdx.x; i < p.S; i += kPhThreads) sm.planes[i] = p.table[i];
    if (threadIdx.x == 0) sm.planes[p.S] = make_uint2(0u, 0xFFFFFFFFu);  // >= 11 mismatches against any valid record
    for (uint32_t i = threadIdx.x; i < nbins; i += kPhThreads) sm.hist[i] = 0u;
    __syncthreads();

    uint64_t *const qh = sm.qhop + warp * kPhQueue;
    uint32_t *const qs = sm.qslow + warp * kPhQueue;
    uint32_t nh = 0u, ns = 0u;  // queue fills, warp-uniform
    const uint64_t total_warps = (uint64_t)gridDim.x * kPhWarps, gwarp = (uint64_t)blockIdx.x * kPhWarps + warp;
    const uint64_t nchunks = (p.n + kPhChunk - 1u) / kPhChunk;

    // records of chunk `ch` into r[0..3]; out-of-range reads become all-ones records (they look invalid, so nothing
    // counts them, and `live` keeps them out of the queues)
    auto fetch = [&](uint64_t ch, uint64_t (&r)[4]) {
        if (ch >= nchunks) return;
        const uint64_t base = ch * kPhChunk;
        if (c.vec && base + kPhChunk <= p.n) {
            const uint64_t *src = c.reads + base + lane * 4u;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]) : "l"(src));
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint64_t i = base + k * 32u + lane;
                r[k] = i < p.n ? __ldg(c.reads + i) : ~0ull;
            }
        }
    };

    uint64_t rec[4], nextrec[4];
    fetch(gwarp, rec);
    uint32_t iter = 0u;
    for (uint64_t ch = gwarp;; ch += total_warps, iter++) {
        const bool final = ch >= nchunks;
        if (!final) {
            fetch(ch + total_warps, nextrec);
            const uint64_t base = ch * kPhChunk;
            const bool packed4 = c.vec && base + kPhChunk <= p.n;  // lane owns 4 consecutive reads
            uint32_t bin[4], dist[4], cnt = 0u, flags = 0u;
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t w0 = (uint32_t)rec[k], w1 = (uint32_t)(rec[k] >> 32);
                const bool slow = ((w0 & c.nm0) | (w1 & c.nm1)) != 0u;
                uint32_t e, d;
                ph_probe(sm, c, w0, w1, e, d);
                const bool ok = d <= p.M && !slow;
                bin[k] = ok ? e : p.S;
                dist[k] = ok ? d : kNoDist;
#ifndef SGDX_PH_X_NOHIST  // (experiment builds only, tools/build_variants.sh: what the histogram costs)
                if (!slow) atomicAdd(&sm.hist[min(dist[k], 2u) * S1 + bin[k]], 1u);
#endif
                const bool live = packed4 || base + k * 32u + lane < p.n;
                const bool hopq = !slow && !ok && p.hop_on != 0u;
                const bool slowq = slow && live;
                cnt += (hopq ? 1u : 0u) + (slowq ? 0x10000u : 0u);
                flags |= ((hopq ? 1u : 0u) | (slowq ? 16u : 0u)) << k;
            }
            // results: a read of the general path gets "Undetermined / no distance" now and is patched by the drain
            if (packed4) {
                const uint64_t at = base + lane * 4u;
                const uint2 v = make_uint2(bin[0] | (bin[1] << 16), bin[2] | (bin[3] << 16));
                *reinterpret_cast<uint2 *>(p.out_idx + at) = v;
                if (p.out_dist)
                    *reinterpret_cast<uint32_t *>(p.out_dist + at) = dist[0] | (dist[1] << 8) | (dist[2] << 16) | (dist[3] << 24);
            } else {
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    const uint64_t i = base + k * 32u + lane;
                    if (i < p.n) {
                        p.out_idx[i] = (uint16_t)bin[k];
                        if (p.out_dist) p.out_dist[i] = (uint8_t)dist[k];
                    }
                }
            }
            // deferred reads -> the warp's queues: one packed inclusive scan gives both write positions
#ifdef SGDX_PH_X_NOQ  // (experiment builds only: what the deferred reads cost)
            cnt = 0u;
#endif
            if (__any_sync(0xFFFFFFFFu, cnt != 0u)) {
                uint32_t incl = cnt;
#pragma unroll
                for (uint32_t o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += t;
                }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                uint32_t ph = nh + ((incl - cnt) & 0xFFFFu), ps = ns + ((incl - cnt) >> 16);
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    if ((flags >> k) & 1u) qh[ph++] = rec[k];
                    if ((flags >> (k + 4u)) & 1u) qs[ps++] = iter * kPhChunk + (packed4 ? lane * 4u + k : k * 32u + lane);
                }
                nh += total & 0xFFFFu;
                ns += total >> 16;
                __syncwarp();
            }
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) rec[k] = nextrec[k];
        }

        // ---- hop check of queued NoMatch reads (metrics.rs:658-676), 32 at a time
        while (nh >= 32u || (final && nh)) {
            const uint32_t take = min(nh, 32u);
            nh -= take;
            if (lane < take) {
                const uint64_t q = qh[nh + lane];
                Rd rd;
                rd.lo = (uint32_t)q & kCompactMask;
                rd.hi = (uint32_t)(q >> kCompactField) & kCompactMask;
                rd.nmask = rd.xmask = 0u;
                hop_check(p, rd);
            }
            __syncwarp();
        }

        // ---- general path: reads with invalid positions or stray bits
        while (ns >= 32u || (final && ns)) {
            const uint32_t take = min(ns, 32u);
            ns -= take;
            bool scan = false;
            uint64_t r = 0;
            if (lane < take) {
                const uint32_t e = qs[ns + lane];
                r = ((uint64_t)(e / kPhChunk) * total_warps + gwarp) * kPhChunk + (e % kPhChunk);
                const Rd rd = rd_of_compact(__ldg(c.reads + r), c.lmask);
                const uint32_t inv = rd.nmask | rd.xmask;
                uint32_t best = kInfKey, nxt = kInfKey;
                if (!cannot_match<MODE>(p, rd)) {
                    if (inv == 0u || (c.single_ok && (inv & (inv - 1u)) == 0u)) {
                        // no invalid base (the record only had stray bits): its own key.  One invalid base: the unique
                        // sample within M over the valid positions, if any, is the entry of one of the 4 completions;
                        // whatever else the probes return is a real sample scored at its real distance
                        const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
                        const uint32_t nvar = inv ? 4u : 1u;
                        for (uint32_t b = 0; b < nvar; b++) {
                            const uint32_t lo = rd.lo | ((b & 1u) ? inv : 0u), hi = rd.hi | ((b >> 1) ? inv : 0u);
                            const uint64_t key = (uint64_t)lo | ((uint64_t)hi << kCompactField);
                            uint32_t pe, pd;
                            ph_probe(sm, c, (uint32_t)key, (uint32_t)(key >> 32), pe, pd);
                            if (pe < p.S) {
                                const uint2 t = sm.planes[pe];
                                score<MODE>(rd, inv, skip, t.x, t.y, pe, best, nxt);
                            }
                        }
                    } else if (p.seed_C == 0u || !seed_search<MODE, true>(p, p.seed_hdr, p.seed_ids, sm.planes, rd, best, nxt)) {
                        scan = true;
                    }
                }
                if (!scan) ph_emit(p, sm.hist, S1, r, rd, apply_rules<MODE>(p, rd, best, nxt));
            }
            // whatever the seed index cannot enumerate: one warp per read over all samples
            uint32_t todo = __ballot_sync(0xFFFFFFFFu, scan);
            while (todo) {
                const uint32_t src = (uint32_t)__ffs(todo) - 1u;
                todo &= todo - 1u;
                const uint64_t qr = __shfl_sync(0xFFFFFFFFu, r, src);
                const Rd qrd = rd_of_compact(__ldg(c.reads + qr), c.lmask);
                const Verdict qv = warp_scan<MODE>(p, sm.planes, qrd, lane);
                if (lane == 0u) ph_emit(p, sm.hist, S1, qr, qrd, qv);
            }
            __syncwarp();
        }
        if (final) break;
    }

    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += kPhThreads) {
        const uint32_t cnt = sm.hist[i];
        if (cnt) {  // hist is [class][bin] with class 0 perfect, 1 one mismatch, 2 other; counters are [bin][other, perfect, one]
            const uint32_t cls = i / S1, b = i % S1;
            atomicAdd(&p.counters[b * 3u + (cls == 2u ? 0u : cls + 1u)], (unsigned long long)cnt);
        }
    }
}

cudaError_t launch_ph(const DevParams &p, const PhParams &c, uint32_t mode, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const size_t smem = ph_smem_bytes(p.S, 32u - c.bshift, 32u - c.sshift);
    auto kern = mode == kModeHamming ? sgdx_match_ph<kModeHamming> : sgdx_match_ph<kModePreCompute>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t chunks = (p.n + kPhChunk - 1u) / kPhChunk, per_cta = kPhWarps;
    uint64_t grid = (chunks + per_cta - 1u) / per_cta;
    if (grid > (uint64_t)sms) grid = (uint64_t)sms;  // one persistent CTA per SM
    kern<<<(uint32_t)grid, kPhThreads, smem, stream>>>(p, c);
    return cudaGetLastError();
}

}  // namespace sgdx
