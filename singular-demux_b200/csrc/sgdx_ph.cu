// sgdx_ph.cu -- sgdx_match_ph: assignment on compact records through a perfect hash of the match neighbourhood (sm_100a).
// This is the kernel of the headline configuration: compact records and a key set that reaches max_mismatches.  Its
// general form -- 16-byte records, key sets of a shorter radius with a second level -- is sgdx_match_phx (sgdx_phx.cu).
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
#include <cstdlib>

#include "sgdx_seed_search.cuh"

namespace sgdx {

constexpr uint32_t kPhThreads = 1024;
constexpr uint32_t kPhWarps = kPhThreads / 32;
constexpr uint32_t kPhChunk = 128;  // reads per warp per iteration
constexpr uint32_t kPhQueue = 160;  // <= 31 left over + <= 128 pushed by one chunk
constexpr uint32_t kPhQueue2 = 64;  // <= 31 left over + <= 32 pushed by one drain of the first queue

struct PhSmem {
    uint64_t *qrec;   // [warps][kPhQueue] deferred reads: record ...
    uint32_t *qoff;   // [warps][kPhQueue] ... and (warp iteration * 128 + offset in chunk)
    uint64_t *q2rec;  // [warps][kPhQueue2] the deferred reads that need the general path ...
    uint32_t *q2off;  // [warps][kPhQueue2] ... same numbering
    uint64_t *hslots; // hop key sets (PhParams::hop_slots), optional
    uint16_t *hdisp;
    uint2 *planes;    // [S + 1][copies]: barcode planes, replicated so that the lanes of a half warp read distinct banks;
                      //                  row S = dummy that nothing matches
    uint32_t *hist;   // [3][S + 1] perfect, one mismatch, other (metrics.rs:428-437), + 1 word nobody reads
    uint16_t *disp;   // [1 << bbits]
    uint16_t *slots;  // [1 << sbits]
};

__host__ __device__ inline size_t ph_carve(unsigned char *base, uint32_t S, uint32_t bbits, uint32_t sbits, uint32_t clog,
                                           uint32_t hsbits, uint32_t hbbits, PhSmem *out) {  // hsbits == 0: no hop tables
    size_t o = 0;
    PhSmem s;
    s.qrec = reinterpret_cast<uint64_t *>(base + o);   o += (size_t)kPhWarps * kPhQueue * 8;
    s.q2rec = reinterpret_cast<uint64_t *>(base + o);  o += (size_t)kPhWarps * kPhQueue2 * 8;
    s.hslots = reinterpret_cast<uint64_t *>(base + o); o += hsbits ? (size_t)8 << hsbits : 0;
    s.planes = reinterpret_cast<uint2 *>(base + o);    o += (((size_t)S + 1) * 8) << clog;
    s.qoff = reinterpret_cast<uint32_t *>(base + o);   o += (size_t)kPhWarps * kPhQueue * 4;
    s.q2off = reinterpret_cast<uint32_t *>(base + o);  o += (size_t)kPhWarps * kPhQueue2 * 4;
    s.hist = reinterpret_cast<uint32_t *>(base + o);   o += ((((size_t)S + 1) * 3 + 1) * 4 + 15) & ~(size_t)15;
    s.disp = reinterpret_cast<uint16_t *>(base + o);   o += (size_t)2 << bbits;
    s.slots = reinterpret_cast<uint16_t *>(base + o);  o += (size_t)2 << sbits;
    s.hdisp = reinterpret_cast<uint16_t *>(base + o);  o += hsbits ? (size_t)2 << hbbits : 0;
    o = (o + 15) & ~(size_t)15;
    if (out) *out = s;
    return o;
}

constexpr size_t kPhSmemLimit = 227u * 1024u;
// A ragged batch (n % 128 != 0) costs a second launch for its last reads, and every launch loads the tables (~14 us),
// while the general kernel is only ~8 % slower per read than the vector one: ragged batches of up to 32 M reads take
// the general kernel alone (measured: 1 M reads 41 -> 27 us, 4 M 49 -> 40 us, 16 M 105 -> 98 us).
constexpr uint64_t kPhOneLaunchBelow = 32u << 20;

// what either kernel needs at least (sgdx_match_phx has a third queue): the host builder sizes its tables against this
size_t ph_smem_bytes(uint32_t S, uint32_t bbits, uint32_t sbits) {
    return ph_carve(nullptr, S, bbits, sbits, 0u, 0u, 0u, nullptr) + (size_t)kPhWarps * kPhQueue2 * 12;
}

// What is optional goes into shared memory as far as it fits: first the hop key sets (without them every hop check
// is a round trip to L2), then copies of the planes table (up to one per lane of a half warp).
struct PhLean {  // per launch: what of the optional tables is in shared memory
    uint32_t clog;      // log2 of the copies of the planes table
    uint32_t hop_smem;  // 1 = the hop key sets too
};

static PhLean ph_fit(uint32_t S, const PhParams &c) {
    PhLean f{0u, 0u};
    const uint32_t bbits = 32u - c.bshift, sbits = 32u - c.sshift;
    const uint32_t hs = 32u - c.hop_sshift, hb = 32u - c.hop_bshift;
    f.hop_smem = c.hop_slots && ph_carve(nullptr, S, bbits, sbits, 0u, hs, hb, nullptr) <= kPhSmemLimit ? 1u : 0u;
    f.clog = 4u;
    while (f.clog && ph_carve(nullptr, S, bbits, sbits, f.clog, f.hop_smem ? hs : 0u, hb, nullptr) > kPhSmemLimit) f.clog--;
    return f;
}

// the perfect-hash probe of one record without invalid positions: entry (sample or S) and distance to it.
// prow = this lane's copy of the planes table, rowbytes = bytes per row of that table.
__device__ __forceinline__ void ph_probe(const PhSmem &sm, const PhParams &c, const unsigned char *prow, uint32_t rowbytes,
                                         uint32_t w0, uint32_t w1, uint32_t &e, uint32_t &d) {
    const uint32_t h = __umulhi(w0, c.k0) + w0 * c.k1 + w1 * c.k2;
    const uint32_t mul = sm.disp[h >> c.bshift];
    e = sm.slots[(h * mul) >> c.sshift];
    const uint2 t = *reinterpret_cast<const uint2 *>(prow + e * rowbytes);
    const uint32_t lo = w0 & kCompactMask;
    const uint32_t hi = __funnelshift_r(w0, w1, kCompactField);  // bits 21..52 of the record; 42..52 are zero here
    d = (uint32_t)__popc((lo ^ t.x) | (hi ^ t.y));
}

// A barcode half in the shared-memory key sets: (index1 id + 1) | (index2 id + 1) << 16, 0 in a field = not such a key.
// key = lo bits | hi bits << length | tag << 40 (PhParams::hop_*)
__device__ __forceinline__ uint32_t ph_hop_find(const PhSmem &sm, const PhParams &c, uint64_t key) {
    const uint32_t w0 = (uint32_t)key, w1 = (uint32_t)(key >> 32);
    const uint32_t h = __umulhi(w0, c.hop_k0) + w0 * c.hop_k1 + w1 * c.hop_k2;
    const uint64_t e = sm.hslots[(h * (uint32_t)sm.hdisp[h >> c.hop_bshift]) >> c.hop_sshift];
    const uint32_t ids = (uint32_t)(e >> 41);  // 11 + 11 bits
    return (e & ((1ull << 41) - 1ull)) == key ? (ids & 0x7FFu) | ((ids >> 11) << 16) : 0u;
}

// the same for a key of at most 32 bits (halves of at most 16 bases): the upper word of the key is zero, so it drops
// out of the hash and the slot's upper word is nothing but the ids
__device__ __forceinline__ uint32_t ph_hop_find32(const PhSmem &sm, const PhParams &c, uint32_t key) {
    const uint32_t h = __umulhi(key, c.hop_k0) + key * c.hop_k1;
    const uint2 e = *reinterpret_cast<const uint2 *>(sm.hslots + ((h * (uint32_t)sm.hdisp[h >> c.hop_bshift]) >> c.hop_sshift));
    const uint32_t ids = e.y >> 9;  // 11 + 11 bits; bits 32..40 of the slot are key bits (zero here)
    return (e.x == key && (e.y & 0x1FFu) == 0u) ? (ids & 0x7FFu) | ((ids >> 11) << 16) : 0u;
}

// metrics.rs:658-676 for an A/C/G/T-only NoMatch read, key sets in shared memory.  Halves of equal length share
// their keys: two look-ups give all four memberships; otherwise four.
__device__ __forceinline__ void ph_hop_check(const DevParams &p, const PhParams &c, const PhSmem &sm, uint32_t lo, uint32_t hi) {
    const uint32_t m1 = (1u << p.i1) - 1u, m2 = (1u << p.i2) - 1u;
    uint32_t a, b, c1, d2;  // ids + 1 of: index1 key in bc[..i1], index2 key in bc[i1..], index1 key in bc[i2..], index2 key in bc[..i2]
    if (p.i1 == p.i2 && p.i1 <= 16u) {
        const uint32_t x = ph_hop_find32(sm, c, (lo & m1) | ((hi & m1) << p.i1));
        const uint32_t y = ph_hop_find32(sm, c, ((lo >> p.i1) & m1) | (((hi >> p.i1) & m1) << p.i1));
        a = x & 0xFFFFu, b = y >> 16, c1 = y & 0xFFFFu, d2 = x >> 16;
    } else if (p.i1 == p.i2) {
        const uint32_t x = ph_hop_find(sm, c, (uint64_t)(lo & m1) | ((uint64_t)(hi & m1) << p.i1));
        const uint32_t y = ph_hop_find(sm, c, (uint64_t)((lo >> p.i1) & m1) | ((uint64_t)((hi >> p.i1) & m1) << p.i1));
        a = x & 0xFFFFu, b = y >> 16, c1 = y & 0xFFFFu, d2 = x >> 16;
    } else {
        const uint64_t tag = 1ull << 40;  // keys of length i2
        a = ph_hop_find(sm, c, (uint64_t)(lo & m1) | ((uint64_t)(hi & m1) << p.i1)) & 0xFFFFu;
        b = ph_hop_find(sm, c, (uint64_t)((lo >> p.i1) & m2) | ((uint64_t)((hi >> p.i1) & m2) << p.i2) | tag) >> 16;
        c1 = ph_hop_find(sm, c, (uint64_t)((lo >> p.i2) & m1) | ((uint64_t)((hi >> p.i2) & m1) << p.i1)) & 0xFFFFu;
        d2 = ph_hop_find(sm, c, (uint64_t)(lo & m2) | ((uint64_t)(hi & m2) << p.i2) | tag) >> 16;
    }
    if (a && b) atomicAdd(&p.hop_fwd[(uint64_t)(a - 1u) * p.u2 + (b - 1u)], 1ull);
    else if (c1 && d2) atomicAdd(&p.hop_rev[(uint64_t)(c1 - 1u) * p.u2 + (d2 - 1u)], 1ull);
}

// verdict of one read of the general path: patch the placeholder the fast path stored and move its count from
// "Undetermined", where the fast path booked it, to its bin; hop check
__device__ __forceinline__ void ph_emit(const DevParams &p, uint32_t *hist, uint32_t S1, uint64_t r, const Rd &rd,
                                        const Verdict &v) {
    if (v.bin != p.S) {
        p.out_idx[r] = (uint16_t)v.bin;
        if (p.out_dist) p.out_dist[r] = (uint8_t)v.dist;
        atomicAdd(&hist[min(v.dist, 2u) * S1 + v.bin], 1u);  // metrics.rs:428-437
        atomicSub(&hist[2u * S1 + p.S], 1u);
    }
    if (p.hop_on && v.bin == p.S) hop_check(p, rd);
}

// VEC: every chunk is full and the three buffers are aligned for 256 / 64 / 32-bit accesses (lane i owns reads
// 4i..4i+3 of a chunk).  !VEC: any alignment, ragged end (lane i owns reads i, i+32, i+64, i+96).
template <uint32_t MODE, bool VEC>
__global__ void __launch_bounds__(kPhThreads, 1) sgdx_match_ph(const DevParams p, const PhParams c, const PhLean f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhSmem sm;
    const uint32_t bbits = 32u - c.bshift, sbits = 32u - c.sshift;
    ph_carve(smem_raw, p.S, bbits, sbits, f.clog, f.hop_smem ? 32u - c.hop_sshift : 0u, 32u - c.hop_bshift, &sm);
    const uint32_t S1 = p.S + 1u, nbins = S1 * 3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ncopies = 1u << f.clog;

    copy_table(sm.disp, c.disp, 2u << bbits, threadIdx.x, kPhThreads);
    copy_table(sm.slots, c.slots, 2u << sbits, threadIdx.x, kPhThreads);
    // (the dummy row is >= 11 mismatches away from any valid record)
    fill_planes(sm.planes, p.table, p.S, f.clog, make_uint2(0u, 0xFFFFFFFFu), threadIdx.x, kPhThreads);
    for (uint32_t i = threadIdx.x; i <= nbins; i += kPhThreads) sm.hist[i] = 0u;
    if (f.hop_smem) {
        copy_table(sm.hslots, c.hop_slots, 8u << (32u - c.hop_sshift), threadIdx.x, kPhThreads);
        copy_table(sm.hdisp, c.hop_disp, 2u << (32u - c.hop_bshift), threadIdx.x, kPhThreads);
    }
    __syncthreads();

    uint64_t *const qrec = sm.qrec + warp * kPhQueue;
    uint32_t *const qoff = sm.qoff + warp * kPhQueue;
    uint64_t *const q2rec = sm.q2rec + warp * kPhQueue2;
    uint32_t *const q2off = sm.q2off + warp * kPhQueue2;
    const unsigned char *const prow = reinterpret_cast<const unsigned char *>(sm.planes + (lane & (ncopies - 1u)));
    const uint32_t rowbytes = 8u << f.clog;
    const uint32_t qrec_s = (uint32_t)__cvta_generic_to_shared(qrec), qoff_s = (uint32_t)__cvta_generic_to_shared(qoff);
    uint32_t n1 = 0u, n2 = 0u;  // queue fills, warp-uniform
    const uint64_t total_warps = (uint64_t)gridDim.x * kPhWarps, gwarp = (uint64_t)blockIdx.x * kPhWarps + warp;
    const uint64_t nchunks = (p.n + kPhChunk - 1u) / kPhChunk;
    const uint32_t below = (1u << lane) - 1u;
    const bool hop = p.hop_on != 0u;
    const uint32_t off0 = VEC ? lane * 4u : lane, offk = VEC ? 1u : 32u;  // chunk offset of my read k = off0 + k * offk

    auto fetch = [&](uint64_t ch, uint64_t (&r)[4]) {
        if (ch >= nchunks) return;
        const uint64_t base = ch * kPhChunk;
        if (VEC) {
            const uint64_t *src = c.reads + base + off0;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]) : "l"(src));
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {  // out-of-range reads become all-ones records: they look invalid
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
            // four reads per lane, phase by phase, so that the four dependent shared-memory chains overlap
            uint32_t bin[4], dist[4], h[4], e[4], defer = 0u;
            uint2 t[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t w0 = (uint32_t)rec[k], w1 = (uint32_t)(rec[k] >> 32);
                h[k] = __umulhi(w0, c.k0) + w0 * c.k1 + w1 * c.k2;
                e[k] = sm.disp[h[k] >> c.bshift];
            }
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) e[k] = sm.slots[(h[k] * e[k]) >> c.sshift];
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) t[k] = *reinterpret_cast<const uint2 *>(prow + e[k] * rowbytes);
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) {
                const uint32_t w0 = (uint32_t)rec[k], w1 = (uint32_t)(rec[k] >> 32);
                const uint32_t lo = w0 & kCompactMask;
                const uint32_t hi = __funnelshift_r(w0, w1, kCompactField);  // bits 21..52; 42..52 are zero unless stray
                const uint32_t d = (uint32_t)__popc((lo ^ t[k].x) | (hi ^ t[k].y));
                const bool stray = ((w0 & c.nm0) | (w1 & c.nm1)) != 0u;  // not "barcode_len A/C/G/T bases"
                const bool bad = d > p.M || stray;
                bin[k] = bad ? p.S : e[k];
                dist[k] = bad ? kNoDist : d;
                // a read of the general path is booked as Undetermined for now; its drain moves the count
                const bool live = VEC || base + k * 32u + lane < p.n;
#ifndef SGDX_PH_X_NOHIST  // (experiment builds only, tools/build_variants.sh: what the histogram costs)
                if (live) atomicAdd(&sm.hist[min(dist[k], 2u) * S1 + bin[k]], 1u);
#endif
                // deferred: the hop check of a NoMatch read, or the whole read if it is not plain A/C/G/T
                if (bad && (hop || stray) && live) defer |= 1u << k;
#ifdef SGDX_PH_X_NOQ  // (experiment builds only: what the deferred reads cost)
                defer = 0u;
#endif
            }
            // results: a read of the general path gets "Undetermined / no distance" now and is patched later
            if (VEC) {
                const uint64_t at = base + off0;
                *reinterpret_cast<uint2 *>(p.out_idx + at) = make_uint2(bin[0] | (bin[1] << 16), bin[2] | (bin[3] << 16));
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
            if (__any_sync(0xFFFFFFFFu, defer != 0u)) {
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) {
                    const bool mine = (defer >> k) & 1u;
                    const uint32_t votes = __ballot_sync(0xFFFFFFFFu, mine);
                    if (mine) {
                        const uint32_t at = n1 + (uint32_t)__popc(votes & below);
                        asm volatile("st.shared.u64 [%0], %1;" ::"r"(qrec_s + at * 8u), "l"(rec[k]) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qoff_s + at * 4u), "r"(iter * kPhChunk + off0 + k * offk) : "memory");
                    }
                    n1 += (uint32_t)__popc(votes);
                }
                __syncwarp();
            }
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) rec[k] = nextrec[k];
        }

        // ---- deferred reads, 32 at a time: hop check of the A/C/G/T-only NoMatch reads (metrics.rs:658-676); the
        // others move on to the second queue
        while (n1 >= 32u || (final && (n1 | n2))) {
            const uint32_t take = min(n1, 32u);  // 0 when only the second queue is left to flush
            n1 -= take;
            bool general = false;
            uint64_t q = 0ull;
            uint32_t off = 0u;
            if (lane < take) {
                q = qrec[n1 + lane];
                off = qoff[n1 + lane];
                const uint32_t w0 = (uint32_t)q, w1 = (uint32_t)(q >> 32);
                general = ((w0 & c.nm0) | (w1 & c.nm1)) != 0u;
                if (!general) {
                    const uint32_t lo = w0 & kCompactMask, hi = __funnelshift_r(w0, w1, kCompactField) & kCompactMask;
                    if (f.hop_smem) {
                        ph_hop_check(p, c, sm, lo, hi);
                    } else {
                        const Rd rd = {lo, hi, 0u, 0u};
                        hop_check(p, rd);
                    }
                }
            }
            const uint32_t votes = __ballot_sync(0xFFFFFFFFu, general);
            if (general) {
                const uint32_t at = n2 + (uint32_t)__popc(votes & below);
                q2rec[at] = q;
                q2off[at] = off;
            }
            n2 += (uint32_t)__popc(votes);
            __syncwarp();

            // ---- general path: reads with invalid positions or stray bits
            while (n2 >= 32u || (final && n1 == 0u && n2)) {
                const uint32_t take2 = min(n2, 32u);
                n2 -= take2;
                bool scan = false;
                uint64_t r = 0;
                if (lane < take2) {
                    const uint32_t e = q2off[n2 + lane];
                    r = ((uint64_t)(e / kPhChunk) * total_warps + gwarp) * kPhChunk + (e % kPhChunk);
                    const Rd rd = rd_of_compact(q2rec[n2 + lane], c.lmask);
                    const uint32_t inv = rd.nmask | rd.xmask;
                    uint32_t best = kInfKey, nxt = kInfKey;
                    if (!cannot_match<MODE>(p, rd)) {
                        if (inv == 0u || (c.single_ok && (inv & (inv - 1u)) == 0u)) {
                            // no invalid base (the record only had stray bits): its own key.  One invalid base: the
                            // unique sample within M over the valid positions, if any, is the entry of one of the 4
                            // completions; whatever else the probes return is a real sample scored at its real distance
                            const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
                            const uint32_t nvar = inv ? 4u : 1u;
                            for (uint32_t b = 0; b < nvar; b++) {
                                const uint32_t lo = rd.lo | ((b & 1u) ? inv : 0u), hi = rd.hi | ((b >> 1) ? inv : 0u);
                                const uint64_t key = (uint64_t)lo | ((uint64_t)hi << kCompactField);
                                uint32_t pe, pd;
                                ph_probe(sm, c, prow, rowbytes, (uint32_t)key, (uint32_t)(key >> 32), pe, pd);
                                if (pe < p.S) {
                                    const uint2 t = *reinterpret_cast<const uint2 *>(prow + pe * rowbytes);
                                    score<MODE>(rd, inv, skip, t.x, t.y, pe, best, nxt);
                                }
                            }
                        } else if (c.single_ok || p.seed_C == 0u ||
                                   !seed_search<MODE, true>(p, p.seed_hdr, p.seed_ids, p.table, rd, best, nxt)) {
                            // with the shortcut valid only reads with several invalid bases get here: too few to be
                            // worth a chain of dependent global loads at one or two active lanes
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
                    const Verdict qv = warp_scan<MODE>(p, p.table, qrd, lane);
                    if (lane == 0u) ph_emit(p, sm.hist, S1, qr, qrd, qv);
                }
                __syncwarp();
            }
        }
        if (final) break;
    }

    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += kPhThreads) {
        const uint32_t cnt = sm.hist[i];
        if (cnt) {  // hist is [class][bin], class 0 perfect, 1 one mismatch, 2 other; counters are [bin][other, perfect, one]
            const uint32_t cls = i / S1, b = i % S1;
            atomicAdd(&p.counters[b * 3u + (cls == 2u ? 0u : cls + 1u)], (unsigned long long)cnt);
        }
    }
}

template <uint32_t MODE, bool VEC>
static cudaError_t launch_ph_t(const DevParams &p, const PhParams &c, const PhLean &f, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const size_t smem = ph_carve(nullptr, p.S, 32u - c.bshift, 32u - c.sshift, f.clog, f.hop_smem ? 32u - c.hop_sshift : 0u,
                                 32u - c.hop_bshift, nullptr);
    auto kern = sgdx_match_ph<MODE, VEC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t chunks = (p.n + kPhChunk - 1u) / kPhChunk;
    uint64_t grid = (chunks + kPhWarps - 1u) / kPhWarps;
    if (grid > (uint64_t)sms) grid = (uint64_t)sms;  // one persistent CTA per SM
    kern<<<(uint32_t)grid, kPhThreads, smem, stream>>>(p, c, f);
    return cudaGetLastError();
}

// Aligned buffers: the full chunks go through the vector kernel and the last n % 128 reads through the general one
// (a second, one-CTA launch); anything else through the general one.
cudaError_t launch_ph(const DevParams &p0, const PhParams &c0, uint32_t mode, int sms, cudaStream_t stream) {
    DevParams p = p0;
    PhParams c = c0;
    const PhLean f = ph_fit(p.S, c);
    const bool ham = mode == kModeHamming;
    const bool aligned = (reinterpret_cast<uintptr_t>(c.reads) & 31u) == 0u && (reinterpret_cast<uintptr_t>(p.out_idx) & 7u) == 0u &&
                         (reinterpret_cast<uintptr_t>(p.out_dist) & 3u) == 0u;
#ifndef SGDX_NO_TEST_HOOKS  // measurement hook: batches of at most this many reads go through the general kernel alone
    static const uint64_t novec_below = getenv("SGDX_PH_NOVEC_BELOW") ? strtoull(getenv("SGDX_PH_NOVEC_BELOW"), nullptr, 10) : kPhOneLaunchBelow;
#else
    static const uint64_t novec_below = kPhOneLaunchBelow;
#endif
    if (aligned && !((p.n & (kPhChunk - 1u)) != 0u && p.n <= novec_below)) {
        const uint64_t n_main = p.n & ~(uint64_t)(kPhChunk - 1u);
        p.n = n_main;
        cudaError_t e = ham ? launch_ph_t<kModeHamming, true>(p, c, f, sms, stream) : launch_ph_t<kModePreCompute, true>(p, c, f, sms, stream);
        if (e != cudaSuccess) return e;
        p.n = p0.n - n_main;
        c.reads += n_main;
        p.out_idx += n_main;
        if (p.out_dist) p.out_dist += n_main;
    }
    return ham ? launch_ph_t<kModeHamming, false>(p, c, f, sms, stream) : launch_ph_t<kModePreCompute, false>(p, c, f, sms, stream);
}

}  // namespace sgdx
