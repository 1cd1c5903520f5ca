// sgdx_exact_first.cuh -- sgdx_match_exact_first: exact-match front end + staged block queues (see the
// description at the top of sgdx_seeded.cu).  The kernel is a template over the number of 32-bit words per
// read; sgdx_exact_words.cu instantiates it once per word count (compiled as eight objects, in parallel).
#pragma once
#include <cstdlib>

#include "sgdx_seed_search.cuh"

namespace sgdx {

// ------------------------------------------------------------------------------------------------
// sgdx_match_exact_first: exact-match front end + queued general path (ASCII input).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kExactReadsPerThread = 4;  // reads per thread between two block barriers

// Work queues of a block (shared memory).  Every stage is run by the whole block on a batch of up to
// blockDim entries, one thread per read, so each stage executes with full warps:
//   front end (all reads)         exact hit -> done            else -> Q1
//   stage 1 (Q1)  neighbour table  one-substitution/one-N hit -> done   else -> Q2
//   stage 2 (Q2)  planes + seed search on reads without invalid bases -> done; reads with invalid bases -> Q3
//   stage 3 (Q3)  seed search with 4-way expansion of an invalid base; windows with several -> warp scan
struct ExactSmem {
    uint2 *table;            // [S] barcode planes (stages 2, 3)
    uint32_t *q1;            // [(R+1)*threads] block-local read numbers
    uint32_t *q2;            // [2*threads]
    uint32_t *q3;            // [2*threads]
    uint32_t *slots;         // [1 << ex_bits]
    uint32_t *ascii;         // [S * Lw]
    uint32_t *hist;          // [(S+1)*3] verdict counters of the block
    uint32_t *cq;            // [threads] Q3 entries that need the all-samples scan
    uint32_t *ring;          // [3] pushes of the current phase, one counter per phase modulo 3
    uint32_t *stage;         // staged front end: the tile of T*R reads a bulk copy delivers (else empty)
    uint64_t *bar;           // ... and its mbarrier
    uint16_t *hdr;           // [seed_hdr_total] seed index
    uint16_t *ids;           // [seed_ids_total]
};

__host__ __device__ inline size_t exact_carve(unsigned char *base, uint32_t threads, uint32_t S, uint32_t Lw,
                                              uint32_t ex_bits, uint32_t hdr_total, uint32_t ids_total,
                                              ExactSmem *out, uint32_t rpt = kExactReadsPerThread,
                                              uint32_t stage_bytes = 0) {
    size_t o = 0;
    ExactSmem s;
    s.bar = reinterpret_cast<uint64_t *>(base + o);          o += stage_bytes ? 128 : 0;                // 8-byte aligned
    s.stage = reinterpret_cast<uint32_t *>(base + o);        o += (stage_bytes + 127u) & ~(size_t)127;  // 16-byte aligned
    s.table = reinterpret_cast<uint2 *>(base + o);           o += (size_t)S * 8;
    s.q1 = reinterpret_cast<uint32_t *>(base + o);           o += (size_t)threads * (rpt + 1) * 4;
    s.q2 = reinterpret_cast<uint32_t *>(base + o);           o += (size_t)threads * 2 * 4;
    s.q3 = reinterpret_cast<uint32_t *>(base + o);           o += (size_t)threads * 2 * 4;
    s.slots = reinterpret_cast<uint32_t *>(base + o);        o += (size_t)4 << ex_bits;
    s.ascii = reinterpret_cast<uint32_t *>(base + o);        o += (size_t)S * Lw * 4;
    s.hist = reinterpret_cast<uint32_t *>(base + o);         o += (size_t)(S + 1) * 3 * 4;
    s.cq = reinterpret_cast<uint32_t *>(base + o);           o += (size_t)threads * 4;
    s.ring = reinterpret_cast<uint32_t *>(base + o);         o += 16;
    s.hdr = reinterpret_cast<uint16_t *>(base + o);          o += ((size_t)hdr_total * 2 + 15) & ~(size_t)15;
    s.ids = reinterpret_cast<uint16_t *>(base + o);          o += ((size_t)ids_total * 2 + 15) & ~(size_t)15;
    if (out) *out = s;
    return o;
}

inline size_t exact_smem_bytes_impl(const DevParams &p, uint32_t threads, uint32_t rpt = kExactReadsPerThread,
                                    uint32_t stage_bytes = 0) {
    return exact_carve(nullptr, threads, p.S, p.Lw, p.ex_bits, p.seed_hdr_total, p.seed_ids_total, nullptr, rpt, stage_bytes);
}

// Staged front end (sgdx_match_exact_first<..., STAGED = true>): the tile of reads arrives in shared memory through
// one 1-D bulk copy (cp.async.bulk + mbarrier) instead of 32-bit loads at the row stride.  Opt-in (SGDX_STAGED=1):
// measured on C2 it does not beat the plain loads yet -- 4 reads per thread need a 40 KB stage (2 CTAs per SM,
// 1.26 ms per 100 M reads against 1.23 ms), 2 reads per thread keep 3 CTAs but double the barriers (1.42 ms).
constexpr uint32_t kStagedReadsPerThread = 4;

__device__ __forceinline__ uint32_t xf_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xf_stage_issue(uint64_t *bar, void *dst, const void *src, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xf_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     xf_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(xf_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void xf_stage_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tXFWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra XFWAIT_%=;\n\t}" ::"r"(xf_smem_u32(bar)), "r"(parity)
        : "memory");
}

// CTA size of sgdx_match_exact_first for this table: 512 threads while at least two CTAs fit an SM (measured on
// C2, ms per 100 M reads: 256 threads 1.51, 512 threads 1.34, 1024 threads 1.48 -- the per-CTA copies of the
// tables are shared by more threads), one 1024-thread CTA per SM for large tables (C3: 1536 x 24 bp needs
// 124 KB of tables), 0 = does not fit at all
inline uint32_t exact_block_threads_impl(const DevParams &p) {
    if (exact_smem_bytes_impl(p, 512) <= 113u * 1024u) return 512u;
    if (exact_smem_bytes_impl(p, 1024) <= 227u * 1024u) return 1024u;
    return 0u;
}

// The same words with the widest loads the row stride allows (base pointer 16-byte aligned, L == 4 * LW): a warp's
// 32 rows span 128 * LW bytes whatever the load width, so every request costs about LW wavefronts in the L1 data
// pipe -- fewer, wider requests mean fewer wavefronts.  Rows of an odd number of words are 8-byte aligned for even
// reads only: odd reads take their first word alone and pair the rest, even reads pair from the start.
template <uint32_t LW>
__device__ __forceinline__ void load_words_wide(const uint8_t *__restrict__ reads, uint64_t r, uint32_t (&rw)[LW]) {
    const uint32_t *pw = reinterpret_cast<const uint32_t *>(reads) + r * LW;
    if constexpr (LW % 4u == 0u) {
#pragma unroll
        for (uint32_t i = 0; i < LW / 4u; i++) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pw) + i);
            rw[4 * i] = v.x, rw[4 * i + 1] = v.y, rw[4 * i + 2] = v.z, rw[4 * i + 3] = v.w;
        }
    } else if constexpr (LW % 2u == 0u) {
#pragma unroll
        for (uint32_t i = 0; i < LW / 2u; i++) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(pw) + i);
            rw[2 * i] = v.x, rw[2 * i + 1] = v.y;
        }
    } else if constexpr (LW == 1u) {
        rw[0] = __ldg(pw);
    } else {
        constexpr uint32_t NP = (LW - 1u) / 2u;
        const uint32_t odd = (uint32_t)r & 1u;
        uint2 v[NP];
#pragma unroll
        for (uint32_t i = 0; i < NP; i++) v[i] = __ldg(reinterpret_cast<const uint2 *>(pw + odd) + i);
        const uint32_t single = __ldg(pw + (odd ? 0u : LW - 1u));
#pragma unroll
        for (uint32_t i = 0; i < NP; i++) {
            rw[2 * i] = odd ? (i == 0u ? single : v[i == 0u ? 0u : i - 1u].y) : v[i].x;
            rw[2 * i + 1] = odd ? v[i].x : v[i].y;
        }
        rw[LW - 1] = odd ? v[NP - 1].y : single;
    }
}

// LW little-endian words of read r; bytes past L read as 'A' (what the table's rows are padded with)
template <uint32_t LW, bool ALIGNED>
__device__ __forceinline__ void load_words(const uint8_t *__restrict__ reads, uint64_t r, uint32_t L,
                                           uint32_t (&rw)[LW]) {
    if (ALIGNED) {
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(reads + r * (uint64_t)(LW * 4u));
#pragma unroll
        for (uint32_t i = 0; i < LW; i++) rw[i] = __ldg(pw + i);
    } else {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(reads) + r * (uint64_t)L;
        const uint32_t a = (uint32_t)(addr & 3u);
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(addr - a);
        const uint32_t k = (a + L + 3u) >> 2;  // aligned words that hold at least one byte of the read
        uint32_t t[LW + 1];
#pragma unroll
        for (uint32_t i = 0; i < LW + 1; i++) t[i] = i < k ? __ldg(pw + i) : 0u;
#pragma unroll
        for (uint32_t i = 0; i < LW; i++) rw[i] = __funnelshift_r(t[i], t[i + 1], a * 8u);
        const uint32_t tail = L & 3u;
        if (tail) {
            const uint32_t keep = (1u << (8u * tail)) - 1u;
            rw[LW - 1] = (rw[LW - 1] & keep) | (0x41414141u & ~keep);
        }
    }
}

template <uint32_t LW>
__device__ __forceinline__ uint32_t word_hash(const DevParams &p, const uint32_t (&rw)[LW]) {
    uint32_t h = 0u;
#pragma unroll
    for (uint32_t i = 0; i < LW; i++) h = ex_word_step(h, rw[i], p.ex_k[i]);
    return ex_mix(h);
}

// planes of a read held as ASCII words (same result as load_ascii)
template <uint32_t LW>
__device__ __forceinline__ Rd planes_of_words(const uint32_t (&rw)[LW], uint32_t L) {
    Rd rd = {0u, 0u, 0u, 0u};
#pragma unroll
    for (uint32_t i = 0; i < LW; i++) fold_word(rw[i], i * 4u, rd);  // pad bytes are 'A': zero bits in every plane
    const uint32_t lm = L >= 32u ? 0xFFFFFFFFu : ((1u << L) - 1u);
    const uint32_t inv = rd.nmask | rd.xmask;
    rd.lo &= lm & ~inv;
    rd.hi &= lm & ~inv;
    return rd;
}

// lo/hi planes of a read held as ASCII words plus a flag "some byte is not A/C/G/T" (then the planes are not
// usable and the caller takes the full conversion).  The expected byte of each 2-bit code comes from a PRMT
// lookup: code = (ascii >> 1) & 3 -> "ACTG"[code]; a byte is valid iff it equals its own code's letter.
template <uint32_t LW>
__device__ __forceinline__ bool planes_if_acgt(const uint32_t (&rw)[LW], uint32_t L, uint32_t &lo, uint32_t &hi) {
    uint32_t l = 0u, h = 0u, bad = 0u;
#pragma unroll
    for (uint32_t i = 0; i < LW; i++) {
        const uint32_t w = rw[i];
        const uint32_t c = (w >> 1) & 0x03030303u;
        const uint32_t y = c | (c >> 4);                               // byte 0: c0 | c1 << 4, byte 2: c2 | c3 << 4
        const uint32_t e = __byte_perm(0x47544341u, 0u, __byte_perm(y, 0u, 0x4420));  // letters of the four codes
        bad |= w ^ e;
        l += (((c & 0x01010101u) * 0x01020408u) >> 24) << (4u * i);    // gather bit 0 of each byte (no carries)
        h += (((c & 0x02020202u) * 0x00810204u) >> 24) << (4u * i);    // gather bit 1 of each byte
    }
    // pad bytes of the last word are 'A' (code 0, valid), so bits >= L are zero already
    lo = l;
    hi = h;
    return bad == 0u;
}

template <uint32_t MODE, uint32_t LW, bool ALIGNED, uint32_t THREADS, bool STAGED = false>
__global__ void __launch_bounds__(THREADS, THREADS == 512u ? 3 : 1) sgdx_match_exact_first(const DevParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t nwarps = THREADS / 32, T = THREADS, R = STAGED ? kStagedReadsPerThread : kExactReadsPerThread;
    constexpr uint32_t kStageBytes = STAGED ? T * R * LW * 4u : 0u;
    ExactSmem sm;
    exact_carve(smem_raw, THREADS, p.S, LW, p.ex_bits, p.seed_hdr_total, p.seed_ids_total, &sm, R, kStageBytes);
    const uint32_t nbins = (p.S + 1u) * 3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    const uint32_t nslots = 1u << p.ex_bits;

    for (uint32_t i = threadIdx.x; i < nslots; i += T) sm.slots[i] = p.ex_slots[i];
    for (uint32_t i = threadIdx.x; i < p.S * LW; i += T) sm.ascii[i] = p.ex_ascii[i];
    for (uint32_t i = threadIdx.x; i < nbins; i += T) sm.hist[i] = 0u;
    for (uint32_t i = threadIdx.x; i < p.S; i += T) sm.table[i] = p.table[i];
    for (uint32_t i = threadIdx.x; i < p.seed_hdr_total; i += T) sm.hdr[i] = p.seed_hdr[i];
    for (uint32_t i = threadIdx.x; i < p.seed_ids_total; i += T) sm.ids[i] = p.seed_ids[i];
    if (threadIdx.x < 4) sm.ring[threadIdx.x] = 0u;
    if (STAGED && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xf_smem_u32(sm.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // staged tiles: full tiles only (their byte count and start are multiples of 16); thread 0 issues the copy of
    // the CTA's next tile as soon as the barrier that ends a front-end pass has freed the stage
    auto stage_next = [&](uint64_t tile_base) {
        if (STAGED && threadIdx.x == 0 && tile_base < p.n && p.n - tile_base >= (uint64_t)(T * R))
            xf_stage_issue(sm.bar, sm.stage, p.reads_ascii + tile_base * (uint64_t)(LW * 4u), kStageBytes);
    };
    stage_next((uint64_t)blockIdx.x * (T * R));
    uint32_t staged_done = 0;

    const uint32_t shift = 32u - p.ex_bits;
    const bool wide = (reinterpret_cast<uintptr_t>(p.reads_ascii) & 15u) == 0u;  // front end: 64/128-bit loads
    uint32_t n1 = 0, n2 = 0, n3 = 0;  // queue fills, tracked identically by every thread

    // queue entries number the block's reads: iteration * (T*R) + offset inside the iteration's tile
    auto read_of = [&](uint32_t e) -> uint64_t {
        return ((uint64_t)blockIdx.x + (uint64_t)(e / (T * R)) * gridDim.x) * (T * R) + (e % (T * R));
    };
    // Every phase (front end or stage) pushes into exactly one queue.  Slots are handed out by an atomic on
    // ring[phase % 3], added to the queue's fill, which every thread tracks in a register.  One barrier ends
    // the phase: it publishes the pushed entries, after it everybody reads the phase's push count, and thread 0
    // clears the counter that phase+2 will use (last read one barrier ago, next written one barrier ahead).
    uint32_t phase = 0;
    auto push_slot = [&]() -> uint32_t { return atomicAdd(&sm.ring[phase % 3u], 1u); };
    auto end_phase = [&]() -> uint32_t {
        __syncthreads();
        const uint32_t pushed = sm.ring[phase % 3u];
        if (threadIdx.x == 0) sm.ring[(phase + 2u) % 3u] = 0u;
        phase++;
        return pushed;
    };

    // ---- stage 3: reads with invalid bases, and whatever the seed index cannot enumerate
    auto stage3 = [&](uint32_t count) {
        const bool valid = threadIdx.x < count;
        uint64_t r = 0;
        Rd rd = {0u, 0u, 0u, 0u};
        uint32_t best = kInfKey, nxt = kInfKey, e = 0;
        bool scan = false;
        if (valid) {
            e = sm.q3[n3 - count + threadIdx.x];
            r = read_of(e);
            uint32_t rw[LW];
            load_words<LW, ALIGNED>(p.reads_ascii, r, p.L, rw);
            rd = planes_of_words<LW>(rw, p.L);
            if (!cannot_match<MODE>(p, rd))
                scan = p.seed_C == 0u || !seed_search<MODE, true>(p, sm.hdr, sm.ids, sm.table, rd, best, nxt);
        }
        if (scan) sm.cq[push_slot()] = e;
        Verdict v = apply_rules<MODE>(p, rd, best, nxt);
        emit(p, r, valid && !scan, rd, v, sm.hist);
        const uint32_t nscan = end_phase();
        n3 -= count;
        if (nscan) {
            for (uint32_t q = warp; q < nscan; q += nwarps) {  // one warp per read, all samples
                const uint64_t qrid = read_of(sm.cq[q]);
                uint32_t rw[LW];
                load_words<LW, ALIGNED>(p.reads_ascii, qrid, p.L, rw);  // same address in every lane
                const Rd qrd = planes_of_words<LW>(rw, p.L);
                Verdict qv = warp_scan<MODE>(p, sm.table, qrd, lane);
                if (lane == 0) emit_single(p, qrid, qrd, qv, sm.hist);
            }
            __syncthreads();  // cq is free again before a following stage 3 pushes into it
        }
    };

    // ---- stage 2: planes + seed search for reads made of A/C/G/T only
    auto stage2 = [&](uint32_t count, const uint32_t *q, uint32_t &nq) {  // q = Q2, or Q1 when there is no stage 1
        const bool valid = threadIdx.x < count;
        uint64_t r = 0;
        Rd rd = {0u, 0u, 0u, 0u};
        uint32_t best = kInfKey, nxt = kInfKey;
        bool done = false;
        if (valid) {
            const uint32_t e = q[nq - count + threadIdx.x];
            r = read_of(e);
            uint32_t rw[LW];
            load_words<LW, ALIGNED>(p.reads_ascii, r, p.L, rw);
            if (!planes_if_acgt<LW>(rw, p.L, rd.lo, rd.hi) || p.seed_C == 0u) {
                sm.q3[n3 + push_slot()] = e;  // invalid bases (stage 3 does the full conversion), or no index
            } else {                          // an all-A/C/G/T read passes the gate and the cannot-match test
                seed_search<MODE, false>(p, sm.hdr, sm.ids, sm.table, rd, best, nxt);
                done = true;
            }
        }
        Verdict v = apply_rules<MODE>(p, rd, best, nxt);
        emit(p, r, done, rd, v, sm.hist);
        n3 += end_phase();
        nq -= count;
    };

    // ---- stage 1: one-substitution / one-no-call neighbours of the barcodes (global table, L2 resident)
    auto stage1 = [&](uint32_t count) {
        const bool valid = threadIdx.x < count;
        if (valid) {
            const uint32_t e = sm.q1[n1 - count + threadIdx.x];
            bool resolved = false;
            {
                const uint64_t r = read_of(e);
                uint32_t rw[LW];
                load_words<LW, ALIGNED>(p.reads_ascii, r, p.L, rw);
                const uint32_t h = word_hash<LW>(p, rw);
                const uint32_t tag = h & 0x7FFu, nshift = 32u - p.nb_bits;
                const uint32_t e1 = __ldg(p.nb_slots + (h >> nshift)), e2 = __ldg(p.nb_slots + (ex_second(h) >> nshift));
                const bool m1 = e1 != kNone && (e1 >> 21) == tag, m2 = e2 != kNone && (e2 >> 21) == tag;
                for (uint32_t t = 0; t < 2u; t++) {  // second round only if both tags matched and the first failed
                    if (!(t == 0u ? (m1 || m2) : (m1 && m2 && !resolved))) continue;
                    const uint32_t ent = t == 0u ? (m1 ? e1 : e2) : e2;
                    const uint32_t s = ent & 0x1FFFu, pos = (ent >> 13) & 31u, letter = (ent >> 18) & 7u;
                    const uint32_t *row = sm.ascii + s * LW;
                    const uint32_t sh = (pos & 3u) * 8u, wi = pos >> 2;
                    const uint32_t byte = letter == 4u ? 0x4Eu : ((0x54474341u >> (8u * letter)) & 0xFFu);  // ACGT, N
                    uint32_t diff = 0u;
#pragma unroll
                    for (uint32_t i = 0; i < LW; i++) {
                        uint32_t expect = row[i];
                        if (i == wi) expect = (expect & ~(0xFFu << sh)) | (byte << sh);
                        diff |= rw[i] ^ expect;
                    }
                    if (diff != 0u) continue;
                    // distance 1 over all positions; every other sample is >= dmin-1 > 1+delta away
                    const bool is_n = letter == 4u;
                    uint32_t d = 1u;
                    bool ok = true;
                    if (is_n) {
                        ok = !(p.max_no_calls >= 0 && 1u > (uint32_t)p.max_no_calls);
                        if (MODE == kModeHamming) d = 1u - min(1u, p.free_ns);
                    }
                    if (ok && d <= p.M) {
                        p.out_idx[r] = (uint16_t)s;
                        if (p.out_dist) p.out_dist[r] = (uint8_t)d;
                        atomicAdd(&sm.hist[s * 3u + (d == 0u ? 1u : 2u)], 1u);
                        resolved = true;
                    }
                    break;  // the read is this neighbour whether or not the rules accept it
                }
            }
            if (!resolved) sm.q2[n2 + push_slot()] = e;
        }
        n2 += end_phase();
        n1 -= count;
    };

    // One loop, one copy of every stage: the deepest non-empty queue with a full batch runs first (so no
    // queue ever exceeds its capacity); when the input is exhausted the queues are flushed the same way.
    uint32_t iter = 0;
    uint64_t base = (uint64_t)blockIdx.x * (T * R);
    for (;;) {
        const bool more = base < p.n;
        if (n3 >= T || (!more && n1 == 0u && n2 == 0u && n3)) {
            stage3(min(n3, T));
        } else if (n2 >= T || (!more && n1 == 0u && n2)) {
            stage2(min(n2, T), sm.q2, n2);
        } else if (n1 >= T || (!more && n1)) {
            if (p.nb_on) stage1(min(n1, T));
            else stage2(min(n1, T), sm.q1, n1);  // no neighbour table: misses go straight to the seed search
        } else if (more) {
            uint32_t rw[R][LW];
            uint32_t hit[R];
            const uint32_t nvalid = (uint32_t)min((uint64_t)(T * R), p.n - base);  // reads of this tile (32-bit tests below)
            bool staged_issued = false;
            if (STAGED && nvalid == T * R) {  // the tile is (being) delivered to shared memory
                xf_stage_wait(sm.bar, staged_done & 1u);
                staged_done++;
#pragma unroll
                for (uint32_t k = 0; k < R; k++)
#pragma unroll
                    for (uint32_t i = 0; i < LW; i++) rw[k][i] = sm.stage[(k * T + threadIdx.x) * LW + i];
                // every thread has its words in registers: refill the stage now, so that the copy of the CTA's
                // next tile runs under this pass's hashing and probing
                __syncthreads();
                stage_next(base + (uint64_t)gridDim.x * (T * R));
                staged_issued = true;
            } else
#pragma unroll
            for (uint32_t k = 0; k < R; k++) {
                const uint64_t r = base + k * T + threadIdx.x;
                if (k * T + threadIdx.x < nvalid) {
                    // measured on a B200: 24-byte rows (C3) 2.01 -> 1.92 ms per 100 M reads with 64-bit loads; rows of
                    // an odd number of words (C2, 20 bytes) lose to the lane-parity selects: 1.24 -> 1.35 ms
                    if (ALIGNED && wide && LW % 2u == 0u) load_words_wide<LW>(p.reads_ascii, r, rw[k]);
                    else load_words<LW, ALIGNED>(p.reads_ascii, r, p.L, rw[k]);
                } else {
#pragma unroll
                    for (uint32_t i = 0; i < LW; i++) rw[k][i] = 0u;  // never equals a barcode row
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < R; k++) {
                const uint64_t r = base + k * T + threadIdx.x;
                const uint32_t h = word_hash<LW>(p, rw[k]);
                const uint32_t tag = h & 0xFFFFu;
                // cuckoo table: the barcode, if this read is one, sits in one of two slots (no probe chains)
                const uint32_t e1 = sm.slots[h >> shift], e2 = sm.slots[ex_second(h) >> shift];
                const bool m1 = e1 != 0u && (e1 >> 16) == tag, m2 = e2 != 0u && (e2 >> 16) == tag;
                hit[k] = kNone;
                if (m1 || m2) {
                    const uint32_t s = ((m1 ? e1 : e2) & 0xFFFFu) - 1u;
                    const uint32_t *row = sm.ascii + s * LW;
                    uint32_t diff = 0u;
#pragma unroll
                    for (uint32_t i = 0; i < LW; i++) diff |= rw[k][i] ^ row[i];
                    if (diff == 0u) hit[k] = s;
                }
                if (m1 && m2 && hit[k] == kNone) {  // both tags matched and the first entry was not it (~never)
                    const uint32_t s = (e2 & 0xFFFFu) - 1u;
                    const uint32_t *row = sm.ascii + s * LW;
                    uint32_t diff = 0u;
#pragma unroll
                    for (uint32_t i = 0; i < LW; i++) diff |= rw[k][i] ^ row[i];
                    if (diff == 0u) hit[k] = s;
                }
                if (hit[k] != kNone) {
                    p.out_idx[r] = (uint16_t)hit[k];
                    if (p.out_dist) p.out_dist[r] = 0u;
                    atomicAdd(&sm.hist[hit[k] * 3u + 1u], 1u);  // perfect match (ATOMS.POPC.INC: aggregated in hardware)
                }
            }
            {   // misses -> Q1: one reservation per warp (every lane adding to the same shared counter costs one
                // wavefront per lane); the lanes take their slots from the ballots
                uint32_t votes[R], total = 0u;
#pragma unroll
                for (uint32_t k = 0; k < R; k++) {
                    votes[k] = __ballot_sync(0xFFFFFFFFu, hit[k] == kNone && k * T + threadIdx.x < nvalid);
                    total += (uint32_t)__popc(votes[k]);
                }
                uint32_t slot = 0u;
                if (lane == 0u && total) slot = atomicAdd(&sm.ring[phase % 3u], total);
                slot = __shfl_sync(0xFFFFFFFFu, slot, 0) + n1;
                const uint32_t below = (1u << lane) - 1u;
#pragma unroll
                for (uint32_t k = 0; k < R; k++) {
                    if ((votes[k] >> lane) & 1u)
                        sm.q1[slot + (uint32_t)__popc(votes[k] & below)] = iter * (T * R) + k * T + threadIdx.x;
                    slot += (uint32_t)__popc(votes[k]);
                }
            }
            n1 += end_phase();
            base += (uint64_t)gridDim.x * (T * R);
            iter++;
            if (!staged_issued) stage_next(base);  // a pass that took ordinary loads leaves the stage free
        } else {
            break;
        }
    }

    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += T) {
        uint32_t c = sm.hist[i];
        if (c) atomicAdd(&p.counters[i], (unsigned long long)c);
    }
}

template <uint32_t MODE, uint32_t LW, bool ALIGNED, uint32_t THREADS, bool STAGED = false>
static cudaError_t launch_exact_t(const DevParams &p, int sms, cudaStream_t stream) {
    constexpr uint32_t R = STAGED ? kStagedReadsPerThread : kExactReadsPerThread;
    size_t smem = exact_smem_bytes_impl(p, THREADS, R, STAGED ? THREADS * R * LW * 4u : 0u);
    auto kern = sgdx_match_exact_first<MODE, LW, ALIGNED, THREADS, STAGED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    kern<<<persistent_grid(p.n, THREADS * R, sms, per_sm), THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

// both rule sets, both alignments and both CTA sizes for one word count
template <uint32_t LW>
cudaError_t launch_exact_words(const DevParams &p, uint32_t mode, int sms, cudaStream_t stream) {
    const bool aligned = (p.L & 3u) == 0u && (reinterpret_cast<uintptr_t>(p.reads_ascii) & 3u) == 0u;
    const bool big = exact_block_threads_impl(p) == 1024u;
    // staged front end: 512-thread CTAs, rows of whole words, 16-byte aligned input, and the table + stage must
    // still leave room for three CTAs per SM
#ifndef SGDX_NO_TEST_HOOKS  // (an experiment kept for its parity test and DESIGN.md 9; slower than the default)
    const bool staged = getenv("SGDX_STAGED") && aligned && !big && (reinterpret_cast<uintptr_t>(p.reads_ascii) & 15u) == 0u &&
                        exact_smem_bytes_impl(p, 512, kStagedReadsPerThread, 512u * kStagedReadsPerThread * LW * 4u) <= 113u * 1024u;
#else
    const bool staged = false;
#endif
    if (staged)
        return mode == kModeHamming ? launch_exact_t<kModeHamming, LW, true, 512, true>(p, sms, stream)
                                    : launch_exact_t<kModePreCompute, LW, true, 512, true>(p, sms, stream);
    if (mode == kModeHamming) {
        if (big) return aligned ? launch_exact_t<kModeHamming, LW, true, 1024>(p, sms, stream)
                                : launch_exact_t<kModeHamming, LW, false, 1024>(p, sms, stream);
        return aligned ? launch_exact_t<kModeHamming, LW, true, 512>(p, sms, stream)
                       : launch_exact_t<kModeHamming, LW, false, 512>(p, sms, stream);
    }
    if (big) return aligned ? launch_exact_t<kModePreCompute, LW, true, 1024>(p, sms, stream)
                            : launch_exact_t<kModePreCompute, LW, false, 1024>(p, sms, stream);
    return aligned ? launch_exact_t<kModePreCompute, LW, true, 512>(p, sms, stream)
                   : launch_exact_t<kModePreCompute, LW, false, 512>(p, sms, stream);
}

}  // namespace sgdx
