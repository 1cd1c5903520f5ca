// sgdx_seed_search.cuh -- device pieces shared by the seeded kernels: candidate scoring, the pigeonhole seed
// search, the one-warp-per-read scan and the single-thread emit (see sgdx_seeded.cu for the description).
#pragma once
#include "sgdx_kernels.h"

namespace sgdx {

// ------------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------------
template <uint32_t MODE>
__device__ __forceinline__ void score(const Rd &rd, uint32_t inv, uint32_t skip, uint32_t slo, uint32_t shi, uint32_t s,
                                      uint32_t &best, uint32_t &nxt) {
    uint32_t t, m;
    asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(t) : "r"(rd.lo), "r"(slo), "r"(inv));  // (lo ^ slo) | inv
    asm("lop3.b32 %0, %1, %2, %3, 0xF6;" : "=r"(m) : "r"(t), "r"(rd.hi), "r"(shi));    // t | (hi ^ shi)
    uint32_t d = __popc(m);
    uint32_t key;
    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(key) : "r"(d), "r"(s));
    if (MODE == kModePreCompute) key = (d == skip) ? kInfKey : key;
    if (key != best) {  // the same sample seen through another chunk
        nxt = min(nxt, max(best, key));
        best = min(best, key);
    }
}

// reads no table can match: every sample is further than M away, or the no-call gate (demux.rs:528-531)
template <uint32_t MODE>
__device__ __forceinline__ bool cannot_match(const DevParams &p, const Rd &rd) {
    const uint32_t n_nc = __popc(rd.nmask), n_x = __popc(rd.xmask);
    bool dead = p.max_no_calls >= 0 && n_nc > (uint32_t)p.max_no_calls;
    if (MODE == kModeHamming) dead |= n_x + (n_nc > p.free_ns ? n_nc - p.free_ns : 0u) > p.M;
    else dead |= rd.xmask != 0u || n_nc > p.M;
    return dead;
}

// Enumerate the seed buckets of one read.  Returns false when some key window holds two or more invalid
// bases (best/nxt are then incomplete and the caller hands the read to the all-samples scan).  Loops are
// kept free of early exits so that the lanes of a warp reconverge after every bucket.
// C chunks, desc_of(c) = descriptor of chunk c (DevParams::seed_desc layout).
// PEEL (only without EXPAND): for an index with long keys, whose buckets are mostly empty -- the first two entries of a
// bucket are scored under predicates, so that the lanes of a warp stay together; a loop takes what is left.
template <uint32_t MODE, bool EXPAND, bool PEEL = false, class DescFn>
__device__ __forceinline__ bool seed_search_t(const DevParams &p, uint32_t C, DescFn desc_of, const uint16_t *__restrict__ hdr,
                                              const uint16_t *__restrict__ ids, const uint2 *__restrict__ table,
                                              const Rd &rd, uint32_t &best, uint32_t &nxt, uint32_t tstride) {
    // tstride: distance of consecutive samples' planes in `table`, in entries (replicated tables of sgdx_ph.cu)
    const uint32_t inv = rd.nmask | rd.xmask;
    const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
    bool complete = true;
    for (uint32_t c = 0; c < C; c++) {
        const uint32_t desc = desc_of(c);
        const uint32_t pos = desc & 0xFFu, w = (desc >> 8) & 0xFFu, off = desc >> 16;
        const uint32_t wm = (1u << w) - 1u;
        const uint32_t iv = (inv >> pos) & wm;
        const uint32_t key0 = ((rd.lo >> pos) & wm) | (((rd.hi >> pos) & wm) << w);
        if (!EXPAND) {  // caller guarantees a read without invalid bases: exactly one bucket per chunk
            uint32_t start = hdr[off + key0];
            const uint32_t end = hdr[off + key0 + 1u];
            if (PEEL) {
#pragma unroll
                for (uint32_t k = 0; k < 2u; k++) {
                    const bool on = start + k < end;
                    const uint32_t s = on ? ids[start + k] : 0u;
                    const uint2 t = table[s * tstride];
                    uint32_t b2 = best, n2 = nxt;
                    score<MODE>(rd, inv, skip, t.x, t.y, s, b2, n2);
                    best = on ? b2 : best;
                    nxt = on ? n2 : nxt;
                }
                start += 2u;
            }
#pragma unroll 1  // buckets hold 0-2 samples: an unrolled loop only adds its trip-count dispatch ladder
            for (uint32_t i = start; i < end; i++) {
                const uint32_t s = ids[i];
                const uint2 t = table[s * tstride];
                score<MODE>(rd, inv, skip, t.x, t.y, s, best, nxt);
            }
            continue;
        }
        const bool several = (iv & (iv - 1u)) != 0u;
        complete = complete && !several;
        const uint32_t nvar = several ? 0u : (iv ? 4u : 1u);
        const uint32_t j = iv ? (uint32_t)__ffs(iv) - 1u : 0u;
        for (uint32_t b = 0; b < nvar; b++) {
            const uint32_t key = key0 | ((b & 1u) << j) | ((b >> 1) << (w + j));
            const uint32_t start = hdr[off + key], end = hdr[off + key + 1u];
#pragma unroll 1
            for (uint32_t i = start; i < end; i++) {
                const uint32_t s = ids[i];
                const uint2 t = table[s * tstride];
                score<MODE>(rd, inv, skip, t.x, t.y, s, best, nxt);
            }
        }
    }
    return complete;
}

// the handle's full seed index (T + 1 chunks, DevParams) -- the same enumeration written out for DevParams::seed_*, as
// the ASCII / seeded kernels and sgdx_match_ph were tuned with it
template <uint32_t MODE, bool EXPAND>
__device__ __forceinline__ bool seed_search(const DevParams &p, const uint16_t *__restrict__ hdr,
                                            const uint16_t *__restrict__ ids, const uint2 *__restrict__ table,
                                            const Rd &rd, uint32_t &best, uint32_t &nxt, uint32_t tstride = 1u) {
    const uint32_t inv = rd.nmask | rd.xmask;
    const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
    bool complete = true;
    for (uint32_t c = 0; c < p.seed_C; c++) {
        const uint32_t desc = p.seed_desc[c];
        const uint32_t pos = desc & 0xFFu, w = (desc >> 8) & 0xFFu, off = desc >> 16;
        const uint32_t wm = (1u << w) - 1u;
        const uint32_t iv = (inv >> pos) & wm;
        const uint32_t key0 = ((rd.lo >> pos) & wm) | (((rd.hi >> pos) & wm) << w);
        if (!EXPAND) {  // caller guarantees a read without invalid bases: exactly one bucket per chunk
            const uint32_t start = hdr[off + key0], end = hdr[off + key0 + 1u];
#pragma unroll 1  // buckets hold 0-2 samples: an unrolled loop only adds its trip-count dispatch ladder
            for (uint32_t i = start; i < end; i++) {
                const uint32_t s = ids[i];
                const uint2 t = table[s * tstride];
                score<MODE>(rd, inv, skip, t.x, t.y, s, best, nxt);
            }
            continue;
        }
        const bool several = (iv & (iv - 1u)) != 0u;
        complete = complete && !several;
        const uint32_t nvar = several ? 0u : (iv ? 4u : 1u);
        const uint32_t j = iv ? (uint32_t)__ffs(iv) - 1u : 0u;
        for (uint32_t b = 0; b < nvar; b++) {
            const uint32_t key = key0 | ((b & 1u) << j) | ((b >> 1) << (w + j));
            const uint32_t start = hdr[off + key], end = hdr[off + key + 1u];
#pragma unroll 1
            for (uint32_t i = start; i < end; i++) {
                const uint32_t s = ids[i];
                const uint2 t = table[s * tstride];
                score<MODE>(rd, inv, skip, t.x, t.y, s, best, nxt);
            }
        }
    }
    return complete;
}

// one warp, one read, all samples
template <uint32_t MODE>
__device__ __forceinline__ Verdict warp_scan(const DevParams &p, const uint2 *__restrict__ table, const Rd &rd,
                                             uint32_t lane) {
    const uint32_t inv = rd.nmask | rd.xmask;
    const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
    uint32_t b = kInfKey, n2 = kInfKey;
    for (uint32_t s0 = 0; s0 < p.S; s0 += 32u) {  // same trip count in every lane: the warp stays converged
        const uint32_t s = s0 + lane;
        if (s < p.S) {
            const uint2 t = table[s];
            score<MODE>(rd, inv, skip, t.x, t.y, s, b, n2);
        }
    }
    const uint32_t gb = __reduce_min_sync(0xFFFFFFFFu, b);
    const uint32_t gn = __reduce_min_sync(0xFFFFFFFFu, b == gb ? n2 : b);  // keys are unique per sample
    return apply_rules<MODE>(p, rd, gb, gn);
}

// verdict of one read written by one thread (the warp-cooperative path)
__device__ __forceinline__ void emit_single(const DevParams &p, uint64_t r, const Rd &rd, const Verdict &v,
                                            uint32_t *hist) {
    p.out_idx[r] = (uint16_t)v.bin;
    if (p.out_dist) p.out_dist[r] = (uint8_t)v.dist;
    uint32_t cls = v.dist == 0u ? 1u : (v.dist == 1u ? 2u : 0u);
    atomicAdd(&hist[v.bin * 3u + cls], 1u);
    if (p.hop_on && v.bin == p.S) hop_check(p, rd);
}

// compact record (include/sgdx.h) <-> planes.  lmask = (1 << barcode_len) - 1: bits a caller left above the
// barcode's length in any plane, and bit 63, are ignored.
__device__ __forceinline__ Rd rd_of_compact(uint64_t rec, uint32_t lmask) {
    const uint32_t lo = (uint32_t)rec & lmask;
    const uint32_t hi = (uint32_t)(rec >> kCompactField) & lmask;
    const uint32_t inv = (uint32_t)(rec >> (2u * kCompactField)) & lmask;
    Rd rd;
    rd.nmask = inv & ~lo;
    rd.xmask = inv & lo;
    rd.lo = lo & ~inv;
    rd.hi = hi & ~inv;
    return rd;
}

__device__ __forceinline__ uint64_t compact_of_rd(const Rd &rd) {
    const uint32_t inv = rd.nmask | rd.xmask;
    return (uint64_t)((rd.lo & ~inv) | rd.xmask) | ((uint64_t)(rd.hi & ~inv) << kCompactField) |
           ((uint64_t)inv << (2u * kCompactField));
}

// Table prologue of the perfect-hash kernels: global -> shared copies in 16-byte pieces (both sides 16-byte aligned;
// a tail of less than 16 bytes goes byte-wise), and the planes table written `copies` times per sample (copies = a
// power of two).  The u16-at-a-time loops they replace were 40 % of a 1 M-read launch.
__device__ __forceinline__ void copy_table(void *dst_smem, const void *src_global, uint32_t bytes, uint32_t tid, uint32_t nthreads) {
    const uint32_t vec = bytes >> 4;
    for (uint32_t i = tid; i < vec; i += nthreads) static_cast<uint4 *>(dst_smem)[i] = __ldg(static_cast<const uint4 *>(src_global) + i);
    for (uint32_t i = (vec << 4) + tid; i < bytes; i += nthreads)
        static_cast<unsigned char *>(dst_smem)[i] = __ldg(static_cast<const unsigned char *>(src_global) + i);
}

__device__ __forceinline__ void fill_planes(uint2 *dst_smem, const uint2 *__restrict__ table, uint32_t S, uint32_t clog, uint2 dummy,
                                            uint32_t tid, uint32_t nthreads) {
    const uint32_t copies = 1u << clog;
    for (uint32_t s = tid; s <= S; s += nthreads) {  // row S = the dummy row
        const uint2 t = s < S ? __ldg(table + s) : dummy;
        uint2 *row = dst_smem + ((size_t)s << clog);
        if (copies >= 2u) {
            for (uint32_t k = 0; k < copies / 2u; k++) reinterpret_cast<uint4 *>(row)[k] = make_uint4(t.x, t.y, t.x, t.y);
        } else {
            row[0] = t;
        }
    }
}

inline uint32_t persistent_grid(uint64_t n, uint32_t threads, int sms, int per_sm) {
    uint64_t tiles = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)sms * (uint64_t)per_sm;  // persistent grid: a multiple of the SM count
    return (uint32_t)(tiles < cap ? (tiles ? tiles : 1) : cap);
}

}  // namespace sgdx
