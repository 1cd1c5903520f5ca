// sgdx_seeded.cu -- the seed-and-verify assignment kernels (sm_100a).
//
// Same results as sgdx_match_direct (matcher.rs:273-348 / 121-262, demux.rs:524-538, metrics.rs:428-437,
// 658-676), but instead of scoring every (read, sample) pair they enumerate only the samples that can
// influence the verdict.
//
// Seed index.  A sample matters only if its number of mismatches over the read's valid (non-N,
// non-foreign) positions is <= T, with T = max_mismatches + min_delta for the Hamming rule set and
// T = max(M, M+delta-1) for the PreCompute rule set ("why T is enough").  Cut the barcode
// into T+1 disjoint chunks: by pigeonhole such a sample agrees with the read on every valid position of
// at least one chunk, so it sits in the bucket that chunk's bases select.
//   - chunk window without invalid positions      -> one bucket
//   - chunk window with exactly one invalid base  -> the 4 buckets of its possible values
//   - anything worse (rare), or no index at all   -> the read is scored against all samples by a whole
//                                                    warp (per-lane best/second-best keys, then
//                                                    __reduce_min_sync across the warp)
// Reads that cannot match whatever the table holds (too many no-calls / foreign bytes for M, or the
// max_no_calls gate) skip the search.  Candidates are scored exactly like the direct kernel:
// popc(((lo^slo)|inv)|(hi^shi)), key = distance << 16 | sample; a sample reached through several
// chunks produces the same key twice, which the `key != best` guard makes harmless.
//
// Exact-match front end (sgdx_match_exact_first).  Most reads of a real run equal their sample's
// barcode byte for byte.  Such a read is Match(sample, 0) without any search whenever the table's
// minimum pairwise distance exceeds min_delta (every other sample is then further than delta away, by
// the triangle inequality), so the kernel first looks the raw ASCII words up in an open-addressing
// table of the barcodes (multiplicative word hash on the fma pipe, 16-bit tag, byte-exact verify).
// Everything else goes to a block-level queue that is drained by the full block, one thread per read,
// through the general path above -- so the divergent work runs with full warps, not with the one or
// two slow lanes a warp of fresh reads would have.
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
template <uint32_t MODE, bool EXPAND>
__device__ __forceinline__ bool seed_search(const DevParams &p, const uint16_t *__restrict__ hdr,
                                            const uint16_t *__restrict__ ids, const uint2 *__restrict__ table,
                                            const Rd &rd, uint32_t &best, uint32_t &nxt) {
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
                const uint2 t = table[s];
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
                const uint2 t = table[s];
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

static uint32_t persistent_grid(uint64_t n, uint32_t threads, int sms, int per_sm) {
    uint64_t tiles = (n + threads - 1) / threads;
    uint64_t cap = (uint64_t)sms * (uint64_t)per_sm;  // persistent grid: a multiple of the SM count
    return (uint32_t)(tiles < cap ? (tiles ? tiles : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// sgdx_match_seeded: one thread per read through the general path, tables in shared memory.
// Used for packed input and for tables whose minimum pairwise distance does not allow the exact-match
// shortcut.
// ------------------------------------------------------------------------------------------------
struct SeedSmem {
    uint4 *dq_rd;        // [threads] planes of the reads queued for the all-samples scan
    uint2 *table;        // [S]
    uint32_t *hist;      // [(S+1)*3]
    uint32_t *dq_tid;    // [threads]
    uint32_t *dq_count;  // [1] (+3 pad)
    uint16_t *hdr;       // [hdr_total]
    uint16_t *ids;       // [ids_total]
};

__host__ __device__ inline size_t seed_carve(unsigned char *base, uint32_t threads, uint32_t S, uint32_t hdr_total,
                                             uint32_t ids_total, SeedSmem *out) {
    size_t o = 0;
    SeedSmem s;
    s.dq_rd = reinterpret_cast<uint4 *>(base + o);      o += (size_t)threads * 16;
    s.table = reinterpret_cast<uint2 *>(base + o);      o += (size_t)S * 8;
    s.hist = reinterpret_cast<uint32_t *>(base + o);    o += (size_t)(S + 1) * 3 * 4;
    s.dq_tid = reinterpret_cast<uint32_t *>(base + o);  o += (size_t)threads * 4;
    s.dq_count = reinterpret_cast<uint32_t *>(base + o); o += 16;
    s.hdr = reinterpret_cast<uint16_t *>(base + o);     o += ((size_t)hdr_total * 2 + 15) & ~(size_t)15;
    s.ids = reinterpret_cast<uint16_t *>(base + o);     o += ((size_t)ids_total * 2 + 15) & ~(size_t)15;
    if (out) *out = s;
    return o;
}

size_t seeded_smem_bytes(const DevParams &p) {
    return seed_carve(nullptr, kSeededThreads, p.S, p.seed_hdr_total, p.seed_ids_total, nullptr);
}

template <uint32_t MODE, bool PACKED>
__global__ void __launch_bounds__(kSeededThreads) sgdx_match_seeded(const DevParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SeedSmem sm;
    seed_carve(smem_raw, kSeededThreads, p.S, p.seed_hdr_total, p.seed_ids_total, &sm);
    const uint32_t nbins = (p.S + 1u) * 3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = kSeededThreads / 32;

    for (uint32_t i = threadIdx.x; i < p.S; i += blockDim.x) sm.table[i] = p.table[i];
    for (uint32_t i = threadIdx.x; i < p.seed_hdr_total; i += blockDim.x) sm.hdr[i] = p.seed_hdr[i];
    for (uint32_t i = threadIdx.x; i < p.seed_ids_total; i += blockDim.x) sm.ids[i] = p.seed_ids[i];
    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sm.hist[i] = 0u;
    if (threadIdx.x == 0) *sm.dq_count = 0u;
    __syncthreads();

    const bool words_ok = (p.L & 3u) == 0u && (reinterpret_cast<uintptr_t>(p.reads_ascii) & 3u) == 0u;

    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = base + threadIdx.x;
        const bool active = r < p.n;
        Rd rd = {0u, 0u, 0u, 0u};
        if (active) rd = PACKED ? load_packed(p.reads_packed, r, p.L) : load_ascii(p.reads_ascii, r, p.L, words_ok);

        uint32_t best = kInfKey, nxt = kInfKey;
        bool queued = false;
        if (active && !cannot_match<MODE>(p, rd)) queued = !seed_search<MODE, true>(p, sm.hdr, sm.ids, sm.table, rd, best, nxt);
        if (queued) {
            uint32_t slot = atomicAdd(sm.dq_count, 1u);
            sm.dq_rd[slot] = make_uint4(rd.lo, rd.hi, rd.nmask, rd.xmask);
            sm.dq_tid[slot] = threadIdx.x;
        }
        Verdict v = apply_rules<MODE>(p, rd, best, nxt);
        emit(p, r, active && !queued, rd, v, sm.hist);

        // queued reads: one warp per read, all samples (rare; the branch is uniform across the block)
        const uint32_t nq = __syncthreads_count(queued);
        if (nq) {
            if (threadIdx.x == 0) *sm.dq_count = 0u;  // every push of this tile happened before the barrier
            for (uint32_t q = warp; q < nq; q += nwarps) {
                const uint4 qv = sm.dq_rd[q];
                const Rd qrd = {qv.x, qv.y, qv.z, qv.w};
                Verdict qvd = warp_scan<MODE>(p, sm.table, qrd, lane);
                if (lane == 0) emit_single(p, base + sm.dq_tid[q], qrd, qvd, sm.hist);
            }
            __syncthreads();  // the queue and its counter are free again before the next tile pushes
        }
    }
    __syncthreads();

    for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) {
        uint32_t c = sm.hist[i];
        if (c) atomicAdd(&p.counters[i], (unsigned long long)c);
    }
}

template <uint32_t MODE, bool PACKED>
static cudaError_t launch_seeded_t(const DevParams &p, int sms, cudaStream_t stream) {
    size_t smem = seeded_smem_bytes(p);
    auto kern = sgdx_match_seeded<MODE, PACKED>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSeededThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    kern<<<persistent_grid(p.n, kSeededThreads, sms, per_sm), kSeededThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

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
    uint16_t *hdr;           // [seed_hdr_total] seed index
    uint16_t *ids;           // [seed_ids_total]
};

__host__ __device__ inline size_t exact_carve(unsigned char *base, uint32_t threads, uint32_t S, uint32_t Lw,
                                              uint32_t ex_bits, uint32_t hdr_total, uint32_t ids_total,
                                              ExactSmem *out) {
    size_t o = 0;
    ExactSmem s;
    s.table = reinterpret_cast<uint2 *>(base + o);           o += (size_t)S * 8;
    s.q1 = reinterpret_cast<uint32_t *>(base + o);           o += (size_t)threads * (kExactReadsPerThread + 1) * 4;
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

size_t exact_smem_bytes(const DevParams &p, uint32_t threads) {
    return exact_carve(nullptr, threads, p.S, p.Lw, p.ex_bits, p.seed_hdr_total, p.seed_ids_total, nullptr);
}

// CTA size of sgdx_match_exact_first for this table: 512 threads while at least two CTAs fit an SM (measured on
// C2, ms per 100 M reads: 256 threads 1.51, 512 threads 1.34, 1024 threads 1.48 -- the per-CTA copies of the
// tables are shared by more threads), one 1024-thread CTA per SM for large tables (C3: 1536 x 24 bp needs
// 124 KB of tables), 0 = does not fit at all
uint32_t exact_block_threads(const DevParams &p) {
    if (exact_smem_bytes(p, 512) <= 113u * 1024u) return 512u;
    if (exact_smem_bytes(p, 1024) <= 227u * 1024u) return 1024u;
    return 0u;
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

template <uint32_t MODE, uint32_t LW, bool ALIGNED, uint32_t THREADS>
__global__ void __launch_bounds__(THREADS) sgdx_match_exact_first(const DevParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExactSmem sm;
    exact_carve(smem_raw, THREADS, p.S, LW, p.ex_bits, p.seed_hdr_total, p.seed_ids_total, &sm);
    const uint32_t nbins = (p.S + 1u) * 3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    constexpr uint32_t nwarps = THREADS / 32, T = THREADS, R = kExactReadsPerThread;
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    const uint32_t nslots = 1u << p.ex_bits;

    for (uint32_t i = threadIdx.x; i < nslots; i += T) sm.slots[i] = p.ex_slots[i];
    for (uint32_t i = threadIdx.x; i < p.S * LW; i += T) sm.ascii[i] = p.ex_ascii[i];
    for (uint32_t i = threadIdx.x; i < nbins; i += T) sm.hist[i] = 0u;
    for (uint32_t i = threadIdx.x; i < p.S; i += T) sm.table[i] = p.table[i];
    for (uint32_t i = threadIdx.x; i < p.seed_hdr_total; i += T) sm.hdr[i] = p.seed_hdr[i];
    for (uint32_t i = threadIdx.x; i < p.seed_ids_total; i += T) sm.ids[i] = p.seed_ids[i];
    if (threadIdx.x < 4) sm.ring[threadIdx.x] = 0u;
    __syncthreads();

    const uint32_t shift = 32u - p.ex_bits;
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
#pragma unroll
            for (uint32_t k = 0; k < R; k++) {
                const uint64_t r = base + k * T + threadIdx.x;
                if (r < p.n) {
                    load_words<LW, ALIGNED>(p.reads_ascii, r, p.L, rw[k]);
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
            {   // misses -> Q1, one slot reservation per thread for all of its reads
                uint32_t nmiss = 0u;
#pragma unroll
                for (uint32_t k = 0; k < R; k++) nmiss += (hit[k] == kNone && base + k * T + threadIdx.x < p.n) ? 1u : 0u;
                if (nmiss) {
                    uint32_t slot = n1 + atomicAdd(&sm.ring[phase % 3u], nmiss);
#pragma unroll
                    for (uint32_t k = 0; k < R; k++)
                        if (hit[k] == kNone && base + k * T + threadIdx.x < p.n)
                            sm.q1[slot++] = iter * (T * R) + k * T + threadIdx.x;
                }
            }
            n1 += end_phase();
            base += (uint64_t)gridDim.x * (T * R);
            iter++;
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

template <uint32_t MODE, uint32_t LW, bool ALIGNED, uint32_t THREADS>
static cudaError_t launch_exact_t(const DevParams &p, int sms, cudaStream_t stream) {
    size_t smem = exact_smem_bytes(p, THREADS);
    auto kern = sgdx_match_exact_first<MODE, LW, ALIGNED, THREADS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    kern<<<persistent_grid(p.n, THREADS * kExactReadsPerThread, sms, per_sm), THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

template <uint32_t MODE, uint32_t LW>
static cudaError_t launch_exact_lw(const DevParams &p, int sms, cudaStream_t stream) {
    if (p.Lw == LW) {
        const bool aligned = (p.L & 3u) == 0u && (reinterpret_cast<uintptr_t>(p.reads_ascii) & 3u) == 0u;
        if (exact_block_threads(p) == 1024u)
            return aligned ? launch_exact_t<MODE, LW, true, 1024>(p, sms, stream)
                           : launch_exact_t<MODE, LW, false, 1024>(p, sms, stream);
        return aligned ? launch_exact_t<MODE, LW, true, 512>(p, sms, stream)
                       : launch_exact_t<MODE, LW, false, 512>(p, sms, stream);
    }
    if constexpr (LW < 8) return launch_exact_lw<MODE, LW + 1>(p, sms, stream);
    return cudaErrorInvalidValue;
}

cudaError_t launch_seeded(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    if (!packed && p.ex_on)
        return mode == kModeHamming ? launch_exact_lw<kModeHamming, 1>(p, sms, stream)
                                    : launch_exact_lw<kModePreCompute, 1>(p, sms, stream);
    if (mode == kModeHamming)
        return packed ? launch_seeded_t<kModeHamming, true>(p, sms, stream)
                      : launch_seeded_t<kModeHamming, false>(p, sms, stream);
    return packed ? launch_seeded_t<kModePreCompute, true>(p, sms, stream)
                  : launch_seeded_t<kModePreCompute, false>(p, sms, stream);
}

}  // namespace sgdx
