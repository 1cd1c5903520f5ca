// sgdx_post.cu -- the stages either side of the assignment kernels, hand-written for sm_100a:
//
//   unmatched table   UnmatchedCounter (metrics.rs:131-190) as a device hash table: 3 bits per base, 128-bit
//                     compare-and-swap claims (ATOMG.CAS.128), threshold selection of the top-N from a
//                     device-side histogram of the counts.
//   sgdx_extract      sample-barcode segments of every FASTQ's read concatenated into the L-byte barcode
//                     (demux.rs:360-417,504-522).
//   sgdx_qual_counts  BaseQualCounter::update (metrics.rs:499-518) keyed by the sample-index array
//                     (demux.rs:582-583).  HBM-bound: tiles of 128 quality strings arrive in shared memory
//                     through 1-D bulk copies (cp.async.bulk + mbarrier, two stages per CTA), one thread per
//                     read counts with byte-lane SWAR arithmetic.
//   sgdx_route        stable partition of the read ids by sample index (what DemuxedGroup::per_sample_reads
//                     holds, demux.rs:637-645): per-warp histograms, bin-major scan, per-warp ranked scatter.
#include <algorithm>
#include <cstring>

#include "sgdx_guard.h"
#include "sgdx_handle.h"
#include "sgdx_seed_search.cuh"  // rd_of_compact

using namespace sgdx;

namespace sgdx {

static uint32_t grid_cap(uint64_t items, uint32_t per_block, int sms, int blocks_per_sm) {
    uint64_t tiles = (items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)sms * (uint64_t)blocks_per_sm;
    return (uint32_t)(tiles < cap ? (tiles ? tiles : 1) : cap);
}

// ================================================================================================
// 1. unmatched barcodes
// ================================================================================================
constexpr int kUnmStats = 8;  // [0] distinct keys in the table  [1] NoMatch reads seen  [2] dropped reads
                              // [3] entries appended to the foreign-byte list  [4] selected entries  [5] tie tickets
constexpr unsigned long long kEmptyWord = ~0ull;
constexpr int kUnmHistBins = 4096;
constexpr unsigned int kUnmPublish = 256;

struct __align__(32) UnmSlot {
    ulonglong2 key;            // 96-bit string, base j = 3 bits at 3j (A0 C1 T2 G3 N7): {bits 0..63, bits 64..95 | 1 << 63};
                               // {~0, ~0} = empty
    unsigned long long count;  // same 32-byte sector as the key: the add after a claim is an L2 hit
    unsigned long long pad;    // 0xFF.. (never read)
};

struct UnmTable {
    UnmSlot *slots;
    unsigned long long *stats;
    uint64_t mask;       // slots - 1
    uint64_t key_limit;  // no new keys once stats[0] reaches this
    uint8_t *odd;        // 32-byte records
    uint64_t odd_cap;
};

__device__ __forceinline__ ulonglong2 cas128(ulonglong2 *addr, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old.x), "=l"(old.y)
        : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
        : "memory");
    return old;
}

__device__ __forceinline__ uint64_t unm_hash(uint64_t k0, uint64_t k1) { return mix64(k0 ^ mix64(k1 + 0x9E3779B97F4A7C15ull)); }

// count[key] += add; a new key claims the first empty slot of its probe sequence with one 128-bit CAS (the CAS is
// also the read: a plain 128-bit load may be torn against a concurrent claim).  `full`: the table may not take new
// keys (then the CAS compares and swaps empty for empty: an atomic read).  0 = not counted (table full),
// 1 = counted, 2 = counted and the key is new.
__device__ __forceinline__ uint32_t unm_add(const UnmTable &t, uint64_t k0, uint64_t k1, unsigned long long add,
                                            bool full) {
    const ulonglong2 empty = make_ulonglong2(kEmptyWord, kEmptyWord), key = make_ulonglong2(k0, k1);
    uint64_t i = unm_hash(k0, k1) & t.mask;
    for (uint64_t probe = 0; probe <= t.mask; probe++) {
        const ulonglong2 old = cas128(&t.slots[i].key, empty, full ? empty : key);
        if (old.x == kEmptyWord && old.y == kEmptyWord) {
            if (full) return 0u;  // the key is not in the table (no deletions) and may not enter it
            atomicAdd(&t.slots[i].count, add + 1ull);  // an empty slot's count is ~0 (memset 0xFF)
            return 2u;
        }
        if (old.x == k0 && old.y == k1) {
            atomicAdd(&t.slots[i].count, add);
            return 1u;
        }
        i = (i + 1) & t.mask;
    }
    return 0u;
}

// four ASCII bases -> four 3-bit codes packed into 12 bits (A0 C1 T2 G3 N7, base k at bit 3k); odd |= a byte that
// is none of ACGTN
__device__ __forceinline__ uint32_t key_word(uint32_t w, uint32_t &odd) {
    const uint32_t c = (w >> 1) & 0x03030303u;
    const uint32_t y = c | (c >> 4);
    const uint32_t x = w ^ __byte_perm(0x47544341u, 0u, __byte_perm(y, 0u, 0x4420));  // zero byte <=> "ACTG"[code]
    const uint32_t bad = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
    const uint32_t nfl = zero_bytes(w ^ 0x4E4E4E4Eu);                                  // 0x80 where the byte is 'N'
    odd |= bad & ~nfl;
    const uint32_t c3 = c | (nfl >> 5);
    const uint32_t t = c3 | (c3 >> 5);
    return (t & 0x3Fu) | ((t >> 10) & 0xFC0u);
}

constexpr uint32_t kUnmTile = 2048;  // reads per CTA step
constexpr uint32_t kUnmThreads = 256;

// NoMatch reads (idx == S) are keyed and counted (demux.rs:553-555).  The rare work runs in full warps: every
// thread tests eight index entries of a tile (eight independent loads) and queues the NoMatch ones in shared
// memory; whenever the queue holds a full batch of 256 the CTA drains it, one thread per queued read (word loads,
// SWAR key, one 128-bit CAS + one add).
// INPUT 0: ASCII barcodes.  1: compact 64-bit records, 2: 16-byte sgdx_read records -- the key comes from the planes
// (every plane byte is spread to every third bit through a 256-entry table in shared memory).  A packed record does
// not keep a byte that is none of A/C/G/T/N, so such a barcode cannot be keyed: it is counted as dropped.
template <int INPUT>
__global__ void __launch_bounds__(kUnmThreads) sgdx_unmatched_count(const void *__restrict__ reads_any,
                                                                    const uint16_t *__restrict__ idx, uint64_t n,
                                                                    uint32_t S, uint32_t L, const UnmTable t) {
    const uint8_t *__restrict__ reads = static_cast<const uint8_t *>(reads_any);
    __shared__ uint32_t s_spread[INPUT ? 256 : 1];  // bit j of the index -> bit 3 j
    if (INPUT)
        for (uint32_t i = threadIdx.x; i < 256u; i += kUnmThreads) {
            uint32_t v = 0u;
            for (uint32_t j = 0; j < 8u; j++) v |= ((i >> j) & 1u) << (3u * j);
            s_spread[i] = v;
        }
    // the number of distinct keys is published in steps of kUnmPublish per CTA (one global atomic per step instead
    // of one per key on a single word); the key limit is therefore enforced to within gridDim.x * kUnmPublish keys
    __shared__ unsigned int s_drop, s_new, s_qn;
    __shared__ uint32_t s_queue[kUnmTile + kUnmThreads];  // entry = tile ordinal of this CTA << 11 | read in tile
    if (threadIdx.x == 0) s_drop = s_new = s_qn = 0u;
    __syncthreads();
    const bool words_ok = (L & 3u) == 0u && (reinterpret_cast<uintptr_t>(reads) & 3u) == 0u;
    const uint32_t nw = (L + 3u) >> 2;
    const uint64_t ntiles = (n + kUnmTile - 1) / kUnmTile;
    uint32_t seen = 0;  // thread 0 only

    auto drain = [&](uint32_t from, uint32_t count) {  // queue entries [from, from + count), count <= 256
        if (INPUT && threadIdx.x < count) {
            const uint32_t e = s_queue[from + threadIdx.x];
            const uint64_t r = ((uint64_t)blockIdx.x + (uint64_t)(e >> 11) * gridDim.x) * kUnmTile + (e & 2047u);
            const unsigned long long distinct = *reinterpret_cast<volatile unsigned long long *>(t.stats);
            Rd rd;
            if (INPUT == 1) rd = rd_of_compact(__ldg(static_cast<const uint64_t *>(reads_any) + r), L >= 32u ? 0xFFFFFFFFu : (1u << L) - 1u);
            else rd = load_packed(static_cast<const uint4 *>(reads_any), r, L);
            bool ok = rd.xmask == 0u;
            if (ok) {
                const uint32_t b0 = rd.lo | rd.nmask, b1 = rd.hi | rd.nmask, b2 = rd.nmask;  // N = 7
                unsigned long long sp[4];
#pragma unroll
                for (uint32_t k = 0; k < 4u; k++)  // bases 8k .. 8k+7 -> 24 bits
                    sp[k] = s_spread[(b0 >> (8u * k)) & 0xFFu] | (s_spread[(b1 >> (8u * k)) & 0xFFu] << 1) |
                            (s_spread[(b2 >> (8u * k)) & 0xFFu] << 2);
                const unsigned long long k0 = sp[0] | (sp[1] << 24) | (sp[2] << 48);
                const unsigned long long k1 = (sp[2] >> 16) | (sp[3] << 8);
                const uint32_t how = unm_add(t, k0, (1ull << 63) | k1, 1ull, distinct >= t.key_limit);
                ok = how != 0u;
                if (how == 2u && (atomicAdd(&s_new, 1u) + 1u) % kUnmPublish == 0u)
                    atomicAdd(t.stats + 0, (unsigned long long)kUnmPublish);
            }
            if (!ok) atomicAdd(&s_drop, 1u);
        }
        if (!INPUT && threadIdx.x < count) {
            const uint32_t e = s_queue[from + threadIdx.x];
            const uint64_t r = ((uint64_t)blockIdx.x + (uint64_t)(e >> 11) * gridDim.x) * kUnmTile + (e & 2047u);
            const uint8_t *p = reads + r * (uint64_t)L;
            const unsigned long long distinct = *reinterpret_cast<volatile unsigned long long *>(t.stats);
            uint32_t w[8];
#pragma unroll
            for (uint32_t i = 0; i < 8u; i++) {
                w[i] = 0x41414141u;  // pad with 'A' (code 0)
                if (i < nw) {
                    if (words_ok) w[i] = __ldg(reinterpret_cast<const uint32_t *>(p) + i);
                    else
                        for (uint32_t k = 0; k < 4u && 4u * i + k < L; k++)
                            w[i] = (w[i] & ~(0xFFu << (8u * k))) | ((uint32_t)__ldg(p + 4u * i + k) << (8u * k));
                }
            }
            unsigned long long k0 = 0ull;
            uint32_t k1 = 0u, odd = 0u;
#pragma unroll
            for (uint32_t i = 0; i < 8u; i++) {
                if (i >= nw) break;
                const uint32_t u = key_word(w[i], odd);  // bases 4i .. 4i+3 at bits 12i .. 12i+11 of a 96-bit string
                if (i < 5u) k0 |= (unsigned long long)u << (12u * i);
                else if (i == 5u) {
                    k0 |= (unsigned long long)u << 60;
                    k1 |= u >> 4;
                } else k1 |= u << (12u * i - 64u);
            }
            bool ok;
            if (!odd) {
                const uint32_t how = unm_add(t, k0, (1ull << 63) | k1, 1ull, distinct >= t.key_limit);
                ok = how != 0u;
                if (how == 2u && (atomicAdd(&s_new, 1u) + 1u) % kUnmPublish == 0u)
                    atomicAdd(t.stats + 0, (unsigned long long)kUnmPublish);
            } else {  // a byte outside ACGTN: keep the bytes themselves, the host counts these
                const unsigned long long at = atomicAdd(t.stats + 3, 1ull);
                ok = at < t.odd_cap;
                if (ok)
                    for (uint32_t j = 0; j < 32u; j++) t.odd[at * 32u + j] = j < L ? __ldg(p + j) : (uint8_t)0;
            }
            if (!ok) atomicAdd(&s_drop, 1u);
        }
    };

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const uint64_t first = tile * kUnmTile;
        const uint32_t cnt = (uint32_t)min((uint64_t)kUnmTile, n - first);
        uint32_t v[kUnmTile / kUnmThreads];
#pragma unroll
        for (uint32_t i = 0; i < kUnmTile / kUnmThreads; i++) {
            const uint32_t rl = threadIdx.x + kUnmThreads * i;
            v[i] = rl < cnt ? (uint32_t)__ldg(idx + first + rl) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (uint32_t i = 0; i < kUnmTile / kUnmThreads; i++)
            if (v[i] == S) s_queue[atomicAdd(&s_qn, 1u)] = (it << 11) | (threadIdx.x + kUnmThreads * i);
        __syncthreads();
        uint32_t qn = s_qn;
        const uint32_t pushed = qn;
        while (qn >= kUnmThreads) {  // full batches from the top of the queue
            qn -= kUnmThreads;
            drain(qn, kUnmThreads);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            seen += pushed - qn;
            s_qn = qn;
        }
        __syncthreads();
    }
    const uint32_t rest = s_qn;
    drain(0u, rest);
    __syncthreads();
    if (threadIdx.x == 0) {
        seen += rest;
        if (seen) atomicAdd(t.stats + 1, (unsigned long long)seen);
        if (s_drop) atomicAdd(t.stats + 2, (unsigned long long)s_drop);
        if (s_new % kUnmPublish) atomicAdd(t.stats + 0, (unsigned long long)(s_new % kUnmPublish));
    }
}

// histogram of min(count, kUnmHistBins-1) over the occupied slots
__global__ void __launch_bounds__(256) sgdx_unmatched_hist(const UnmSlot *__restrict__ tab, uint64_t slots,
                                                           unsigned long long *__restrict__ hist) {
    __shared__ unsigned int sh[kUnmHistBins];
    for (uint32_t i = threadIdx.x; i < kUnmHistBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x) {
        if (tab[i].key.y == kEmptyWord) continue;  // occupied slots carry bit 63 only in the top half of .y
        const unsigned long long c = tab[i].count;
        atomicAdd(&sh[c < (unsigned long long)(kUnmHistBins - 1) ? (uint32_t)c : (uint32_t)(kUnmHistBins - 1)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kUnmHistBins; i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// entries with count > t, plus up to `quota` entries with count == t, compacted into out_*
__global__ void __launch_bounds__(256) sgdx_unmatched_select(const UnmSlot *__restrict__ tab,
                                                             uint64_t slots, unsigned long long t,
                                                             unsigned long long quota, ulonglong2 *__restrict__ out_keys,
                                                             unsigned long long *__restrict__ out_counts,
                                                             uint64_t out_cap, unsigned long long *__restrict__ stats) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 k = tab[i].key;
        if (k.y == kEmptyWord) continue;
        const unsigned long long c = tab[i].count;
        bool take = c > t;
        if (c == t && quota) take = atomicAdd(stats + 5, 1ull) < quota;
        if (!take) continue;
        const unsigned long long j = atomicAdd(stats + 4, 1ull);
        if (j < out_cap) {
            out_keys[j] = k;
            out_counts[j] = c;
        }
    }
}

__global__ void __launch_bounds__(256) sgdx_unmatched_reinsert(const ulonglong2 *__restrict__ sel_keys,
                                                               const unsigned long long *__restrict__ sel_counts,
                                                               uint64_t n, const UnmTable t) {
    unsigned int fresh = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        fresh += unm_add(t, sel_keys[i].x, sel_keys[i].y, sel_counts[i], false) == 2u;
    if (fresh) atomicAdd(t.stats + 0, (unsigned long long)fresh);
}

// ================================================================================================
// 2. segment extraction
// ================================================================================================
struct ExtractParams {
    // the sample-barcode segments in concatenation order: segment i copies run_len[i] bytes from
    // run_ptr[i] + read * run_stride[i] (its FASTQ's matrix + the segment's offset in the read)
    const uint8_t *run_ptr[SGDX_MAX_SEGMENTS];
    uint32_t run_stride[SGDX_MAX_SEGMENTS];
    uint32_t run_len[SGDX_MAX_SEGMENTS];
    uint32_t runs, L;
    uint64_t n;
    uint8_t *out;
    uint32_t words_ok;                       // out is 4-byte aligned
};
constexpr uint32_t kExtractThreads = 256;  // = reads per tile

// One thread per read copies the byte runs of its sample-barcode segments (the segment list is indexed by the loop
// counter, so it comes through the uniform datapath; inside a run the addresses are consecutive) into a
// shared-memory tile; the CTA then writes the tile with coalesced 32-bit stores.  A 20-byte row stride is
// conflict-free for the byte stores (bank = 5 * lane + pos / 4).
// OUT 0: the L-byte ASCII barcodes.  1 / 2: the barcode goes straight into a compact 64-bit record / a 16-byte
// sgdx_read record (the thread that gathered a row folds it into planes: BASELINE's "reads packed into 2-bit base
// planes plus an N mask" inside the kernels), so that sgdx_match_segments_device never writes or re-reads the ASCII form.
template <int OUT>
__global__ void __launch_bounds__(kExtractThreads) sgdx_extract(const ExtractParams p) {
    __shared__ __align__(16) uint8_t s_tile[kExtractThreads * SGDX_MAX_BARCODE_LEN];
    const uint32_t L = p.L;
    const uint64_t ntiles = (p.n + kExtractThreads - 1) / kExtractThreads;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t first = tile * kExtractThreads;
        const uint32_t cnt = (uint32_t)min((uint64_t)kExtractThreads, p.n - first);
        if (threadIdx.x < cnt) {
            const uint64_t r = first + threadIdx.x;
            uint8_t *row = s_tile + threadIdx.x * L;
            for (uint32_t g = 0; g < p.runs; g++) {
                const uint8_t *from = p.run_ptr[g] + r * p.run_stride[g];
                const uint32_t len = p.run_len[g];
#pragma unroll 4
                for (uint32_t k = 0; k < len; k++) row[k] = __ldg(from + k);
                row += len;
            }
        }
        if (OUT != 0) {
            if (threadIdx.x < cnt) {
                const uint8_t *row = s_tile + threadIdx.x * L;  // written by this thread
                Rd rd = {0u, 0u, 0u, 0u};
                if ((L & 3u) == 0u) {  // rows start on word boundaries
                    for (uint32_t j = 0; j < L; j += 4) fold_word(*reinterpret_cast<const uint32_t *>(row + j), j, rd);
                } else {
                    for (uint32_t j = 0; j < L; j += 4) {
                        uint32_t w = 0x41414141u;  // pad with 'A': contributes zero bits to every plane
                        for (uint32_t k = 0; k < 4 && j + k < L; k++) w = (w & ~(0xFFu << (8 * k))) | ((uint32_t)row[j + k] << (8 * k));
                        fold_word(w, j, rd);
                    }
                }
                const uint32_t inv = rd.nmask | rd.xmask;
                rd.lo &= ~inv;
                rd.hi &= ~inv;
                if (OUT == 1) reinterpret_cast<uint64_t *>(p.out)[first + threadIdx.x] = compact_of_rd(rd);
                else reinterpret_cast<uint4 *>(p.out)[first + threadIdx.x] = make_uint4(rd.lo, rd.hi, rd.nmask, rd.xmask & ~rd.nmask);
            }
            continue;  // (rows are private to their threads: no barrier needed)
        }
        __syncthreads();
        const uint32_t bytes = cnt * L;
        uint8_t *out = p.out + first * L;  // first * L is a multiple of 4
        if (p.words_ok) {
            const uint32_t words = bytes >> 2;
            for (uint32_t i = threadIdx.x; i < words; i += kExtractThreads)
                reinterpret_cast<uint32_t *>(out)[i] = reinterpret_cast<const uint32_t *>(s_tile)[i];
            for (uint32_t i = (words << 2) + threadIdx.x; i < bytes; i += kExtractThreads) out[i] = s_tile[i];
        } else {
            for (uint32_t i = threadIdx.x; i < bytes; i += kExtractThreads) out[i] = s_tile[i];
        }
        __syncthreads();
    }
}

// ================================================================================================
// 3. base-quality counters
// ================================================================================================
constexpr int kQualThreads = 128;
constexpr int kQualSplit = 1;                            // threads per read (2 was measured: slower)
constexpr int kQualCounters = 5;  // q20, q30, index_bases_qual_sum, index_bases_total_seen, bases (metrics.rs:468-481)

struct QualParams {
    const uint8_t *quals;
    const uint32_t *lens;  // nullable
    const uint16_t *idx;
    uint64_t n;
    uint32_t stride, S;
    uint32_t reads_per_tile;  // multiple of 16
    uint32_t stage_bytes;     // shared-memory bytes of one stage (multiple of 128)
    uint32_t stages;
    uint32_t nwords;          // words per shifted mask
    uint32_t min_len;         // a shorter read cannot be cut into the fixed segments
    uint32_t one;             // 1 (see add_fma)
    uint32_t short_rows;      // stride <= 32: one SWAR pass per row instead of 16-byte chunks
    uint32_t tmask32, bmask32;  // ... with the template / sample-barcode position bits of the structure
    uint32_t t_tail;          // positions t_tail .. stride-1 are all template bases (stride if the last one is not)
    uint32_t smem_hist;       // per-CTA histogram in shared memory (else global atomics)
    const uint32_t *masks;    // [2][16][nwords]: template / sample-barcode position bits, shifted left by 0..15
    unsigned long long *counters;  // [(S+1)][5], then 1 word: short reads
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (both addresses and the size are
// multiples of 16)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// a + k on the fma pipe (IMAD): the alu pipe carries the LOP3 / LEA.HI of the counting loop
// (`one` is a kernel parameter holding 1: a literal multiplier would be folded back into an alu-pipe add)
__device__ __forceinline__ uint32_t add_fma(uint32_t a, uint32_t one, uint32_t k) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(k));
    return r;
}

// x & 0x80808080, opaque to the optimiser so that `acc += top_bits(..) >> 7` stays one LEA.HI
__device__ __forceinline__ uint32_t top_bits(uint32_t x) {
    uint32_t r;
    asm("and.b32 %0, %1, 0x80808080;" : "=r"(r) : "r"(x));
    return r;
}

__device__ __forceinline__ uint32_t lanes8_sum(uint32_t acc) {  // sum of the four byte lanes
    const uint32_t t = (acc & 0x00FF00FFu) + ((acc >> 8) & 0x00FF00FFu);
    return (t & 0xFFFFu) + (t >> 16);
}

constexpr int kQualShortThreads = 256;  // rows of <= 32 bytes: 32 registers, so twice the warps fit

template <bool HAS_B, bool SHORT_ROWS>
__global__ void __launch_bounds__(SHORT_ROWS ? kQualShortThreads : kQualThreads) sgdx_qual_counts(const QualParams p) {
    constexpr uint32_t kThreads = SHORT_ROWS ? kQualShortThreads : kQualThreads, kLanes = kThreads / kQualSplit;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *stage0 = smem;
    uint32_t *s_masks = reinterpret_cast<uint32_t *>(smem + (size_t)p.stages * p.stage_bytes);
    uint32_t *hist = s_masks + 32u * p.nwords;
    const uint32_t nbins = p.smem_hist ? (p.S + 1u) * kQualCounters : 0u;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(hist + ((nbins + 1u) & ~1u));
    const uint32_t tid = threadIdx.x;

    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; s++) mbar_init(mbar + s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < 32u * p.nwords; i += kThreads) s_masks[i] = p.masks[i];
    for (uint32_t i = tid; i < nbins; i += kThreads) hist[i] = 0u;
    __syncthreads();

    const uint32_t R = p.reads_per_tile;
    const uint64_t ntiles = (p.n + R - 1) / R;
    auto issue = [&](uint64_t tile, uint32_t s) {  // thread 0 only
        const uint64_t first = tile * R;
        const uint32_t cnt = (uint32_t)min((uint64_t)R, p.n - first);
        const uint32_t bulk = (cnt * p.stride) & ~15u;
        mbar_expect_tx(mbar + s, bulk);
        if (bulk) bulk_g2s(stage0 + (size_t)s * p.stage_bytes, p.quals + first * p.stride, bulk, mbar + s);
    };
    if (tid == 0)
        for (uint32_t s = 0; s < p.stages; s++) {
            const uint64_t tile = (uint64_t)blockIdx.x + (uint64_t)s * gridDim.x;
            if (tile < ntiles) issue(tile, s);
        }

    unsigned long long *gcnt = p.counters;
    auto bump = [&](uint32_t slot, uint32_t v) {
        if (p.smem_hist) atomicAdd(hist + slot, v);
        else atomicAdd(gcnt + slot, (unsigned long long)v);
    };
    auto flush = [&]() {
        __syncthreads();
        for (uint32_t i = tid; i < nbins; i += kThreads) {
            const uint32_t c = hist[i];
            if (c) {
                atomicAdd(gcnt + i, (unsigned long long)c);
                hist[i] = 0u;
            }
        }
        __syncthreads();
    };

    uint32_t it = 0, shorts = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const uint32_t s = it % p.stages, parity = (it / p.stages) & 1u;
        unsigned char *stage = stage0 + (size_t)s * p.stage_bytes;
        const uint64_t first = tile * R;
        const uint32_t cnt = (uint32_t)min((uint64_t)R, p.n - first);
        const uint32_t bytes = cnt * p.stride, bulk = bytes & ~15u;
        // this thread's first read of the tile: its sample index and length do not depend on the tile's bytes, so
        // their latency is spent while the bulk copy is still in flight
        const uint32_t lane_read = tid / kQualSplit, part = tid % kQualSplit;
        uint32_t bin0 = 0, len0 = p.stride;
        if (lane_read < cnt) {
            bin0 = __ldg(p.idx + first + lane_read);
            if (p.lens) len0 = min(__ldg(p.lens + first + lane_read), p.stride);
        }
        mbar_wait(mbar + s, parity);
        if (bytes != bulk) {  // last tile only: the < 16 bytes a bulk copy cannot carry
            if (tid < bytes - bulk) stage[bulk + tid] = p.quals[first * p.stride + bulk + tid];
            __syncthreads();
        }
        for (uint32_t rl = lane_read; rl < cnt; rl += kLanes) {  // several reads per thread when they are short
            const uint32_t start = rl * p.stride, a = start & 15u;
            const uint4 *chunks = reinterpret_cast<const uint4 *>(stage + (start & ~15u));
            uint32_t len = len0, bin = bin0;
            if (rl != lane_read) {
                bin = __ldg(p.idx + first + rl);
                len = p.lens ? min(__ldg(p.lens + first + rl), p.stride) : p.stride;
            }
            if (len < p.min_len) {
                shorts += part == 0u;
            } else if (SHORT_ROWS && bin <= p.S) {
                // rows of at most 32 bytes (index reads): the chunk machinery costs more than the bytes.  The row's
                // words come from aligned 32-bit loads and a funnel shift, the position masks are two 32-bit
                // kernel parameters, the sums are byte-lane SWAR; a word with a byte outside 33..127 (also a
                // neighbour's or a stale one: only ever a false alarm) sends the row to the exact byte loop.
                const uint32_t vm = len >= 32u ? 0xFFFFFFFFu : ((1u << len) - 1u);
                const uint32_t tmb = p.tmask32 & vm, bmb = HAS_B ? (p.bmask32 & vm) : 0u;
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(stage + (start & ~3u));
                const uint32_t sh = (start & 3u) * 8u, nrw = (p.stride + 3u) >> 2;
                uint32_t a20 = 0, a30 = 0, bacc = 0, flag = 0;
                for (uint32_t i = 0; i < nrw; i++) {
                    const uint32_t x = __funnelshift_r(wp[i], wp[i + 1], sh);  // bytes 4i .. 4i+3 of the row
                    flag |= (x - 0x21212121u) | x;
                    const uint32_t t80 = (((tmb >> (4u * i)) & 15u) * 0x10204080u) & 0x80808080u;
                    a20 += ((x + 0x4B4B4B4Bu) & t80) >> 7;
                    a30 += ((x + 0x41414141u) & t80) >> 7;
                    if (HAS_B) {
                        const uint32_t ff = ((((bmb >> (4u * i)) & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;
                        const uint32_t xb = x & ff;
                        bacc += (xb & 0x00FF00FFu) + ((xb >> 8) & 0x00FF00FFu);
                    }
                }
                uint32_t c20, c30, bsum;
                const uint32_t nT = __popc(tmb), nB = __popc(bmb);
                if ((flag & 0x80808080u) == 0u) {
                    c20 = lanes8_sum(a20);
                    c30 = lanes8_sum(a30);
                    bsum = (bacc & 0xFFFFu) + (bacc >> 16) - 33u * nB;
                } else {
                    c20 = c30 = bsum = 0u;
                    const unsigned char *row = stage + start;
                    for (uint32_t j = 0; j < len; j++) {
                        const uint32_t q = ((uint32_t)row[j] - 33u) & 0xFFu;
                        if ((tmb >> j) & 1u) {
                            c20 += q >= 20u;
                            c30 += q >= 30u;
                        }
                        if ((bmb >> j) & 1u) bsum += q;
                    }
                }
                if (nT) {
                    if (c20) bump(bin * kQualCounters + 0u, c20);
                    if (c30) bump(bin * kQualCounters + 1u, c30);
                    bump(bin * kQualCounters + 4u, nT);
                }
                if (HAS_B && nB) {
                    if (bsum) bump(bin * kQualCounters + 2u, bsum);
                    bump(bin * kQualCounters + 3u, nB);
                }
            } else if (!SHORT_ROWS && bin <= p.S) {
                const uint32_t *tm = s_masks + a * p.nwords, *bm = s_masks + (16u + a) * p.nwords;
                const uint32_t end = a + len, nchunks = (end + 15u) >> 4;
                uint32_t acc20 = 0, acc30 = 0, c20 = 0, c30 = 0, nT = 0;   // template bases
                uint32_t bacc = 0, bsum = 0, nB = 0, nBfast = 0;            // sample-barcode bases
                // byte lanes gain <= 4 per chunk, the 16-bit lanes of bacc <= 1016 per (general) chunk: at most 15 + 32
                // + 15 chunks pass between two drains below
                auto drain = [&]() {
                    c20 += lanes8_sum(acc20);
                    c30 += lanes8_sum(acc30);
                    bsum += (bacc & 0xFFFFu) + (bacc >> 16);
                    acc20 = acc30 = bacc = 0u;
                };
                // a byte outside 33..127 in the chunk: exact u8 arithmetic, byte by byte
                auto slow_chunk = [&](const uint32_t (&x)[4], uint32_t m, uint32_t mb) {
#pragma unroll 1
                    for (uint32_t j = 0; j < 16u; j++) {
                        const uint32_t xw = j < 8u ? (j < 4u ? x[0] : x[1]) : (j < 12u ? x[2] : x[3]);
                        const uint32_t q = (((xw >> (8u * (j & 3u))) & 0xFFu) - 33u) & 0xFFu;
                        if ((m >> j) & 1u) {
                            c20 += q >= 20u;
                            c30 += q >= 30u;
                            nT++;
                        }
                        if ((mb >> j) & 1u) {
                            bsum += q;
                            nB++;
                        }
                    }
                };
                // any chunk: position masks from the read structure, trimmed to the read's own length
                auto general_chunk = [&](uint32_t k) {
                    const uint32_t sh = (k & 1u) * 16u;
                    uint32_t m = (tm[k >> 1] >> sh) & 0xFFFFu;
                    uint32_t mb = HAS_B ? (bm[k >> 1] >> sh) & 0xFFFFu : 0u;
                    const uint32_t rem = end - 16u * k;
                    if (rem < 16u) {
                        const uint32_t lm = (1u << rem) - 1u;
                        m &= lm;
                        mb &= lm;
                    }
                    if ((m | mb) == 0u) return;
                    const uint4 v = chunks[k];
                    const uint32_t x[4] = {v.x, v.y, v.z, v.w};
                    // every byte in 33..127 <=> no 0x80 bit in (x - 0x21..) | x (a borrow only starts at an
                    // offending byte, which flags itself)
                    uint32_t flag = 0;
#pragma unroll
                    for (int w = 0; w < 4; w++) flag |= (x[w] - 0x21212121u) | x[w];
                    if (flag & 0x80808080u) {
                        slow_chunk(x, m, mb);
                        return;
                    }
                    if (m) {  // q >= 20 <=> byte + (128 - 53) carries into bit 7; no carry leaves a byte
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            const uint32_t t80 = (((m >> (4 * w)) & 15u) * 0x10204080u) & 0x80808080u;
                            acc20 += ((x[w] + 0x4B4B4B4Bu) & t80) >> 7;
                            acc30 += ((x[w] + 0x41414141u) & t80) >> 7;
                        }
                        nT += __popc(m);
                    }
                    if (HAS_B && mb) {
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            const uint32_t ff = ((((mb >> (4 * w)) & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;
                            const uint32_t xb = x[w] & ff;
                            bacc += (xb & 0x00FF00FFu) + ((xb >> 8) & 0x00FF00FFu);
                        }
                        nBfast += __popc(mb);
                    }
                };
                // chunks [k_lo, k_hi) hold template bases only (the structure ends in a run of T from position
                // t_tail on, and the chunk lies inside the read): no masks
                const uint32_t k_lo = min((p.t_tail + a + 15u) >> 4, nchunks), k_hi = max(end >> 4, k_lo);
                // this thread's share of the read: chunks [kb, ke)
                const uint32_t kb = (nchunks * part) / kQualSplit, ke = (nchunks * (part + 1u)) / kQualSplit;
                const uint32_t g0 = min(k_lo, ke), f0 = max(k_lo, kb), f1 = max(min(k_hi, ke), f0), g1 = max(k_hi, kb);
                for (uint32_t k = kb; k < g0; k++) {
                    general_chunk(k);
                    if (((k - kb) & 15u) == 15u) drain();
                }
                for (uint32_t k0 = f0; k0 < f1; k0 += 32u) {
                    // branch-free over the block; a byte outside 33..127 anywhere in it (never in a valid FASTQ)
                    // discards the block's sums and redoes it chunk by chunk
                    const uint32_t k1 = min(k0 + 32u, f1);
                    const uint32_t save20 = acc20, save30 = acc30;
                    uint32_t flag = 0;
#pragma unroll 4
                    for (uint32_t k = k0; k < k1; k++) {
                        const uint4 v = chunks[k];
                        const uint32_t x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            flag |= (x[w] - 0x21212121u) | x[w];
                            acc20 += top_bits(add_fma(x[w], p.one, 0x4B4B4B4Bu)) >> 7;  // IMAD, LOP3, LEA.HI
                            acc30 += top_bits(add_fma(x[w], p.one, 0x41414141u)) >> 7;
                        }
                    }
                    if (flag & 0x80808080u) {
                        acc20 = save20;
                        acc30 = save30;
                        for (uint32_t k = k0; k < k1; k++) {
                            general_chunk(k);
                            if (((k - k0) & 15u) == 15u) drain();
                        }
                        drain();
                    } else {
                        nT += 16u * (k1 - k0);
                        if (k1 < f1) drain();
                    }
                }
                for (uint32_t k = g1; k < ke; k++) {
                    general_chunk(k);
                    if (((k - g1) & 15u) == 15u) drain();
                }
                c20 += lanes8_sum(acc20);
                c30 += lanes8_sum(acc30);
                if (nT) {
                    if (c20) bump(bin * kQualCounters + 0u, c20);
                    if (c30) bump(bin * kQualCounters + 1u, c30);
                    bump(bin * kQualCounters + 4u, nT);
                }
                if (HAS_B) {
                    bsum += (bacc & 0xFFFFu) + (bacc >> 16) - 33u * nBfast;
                    nB += nBfast;
                    if (nB) {
                        if (bsum) bump(bin * kQualCounters + 2u, bsum);
                        bump(bin * kQualCounters + 3u, nB);
                    }
                }
            }
        }
        __syncthreads();  // every thread is done with this stage before it is refilled
        if (tid == 0) {
            const uint64_t next = tile + (uint64_t)p.stages * gridDim.x;
            if (next < ntiles) issue(next, s);
        }
        if (p.smem_hist && (it & 63u) == 63u) flush();  // 32-bit bins: 64 tiles of <= 48 KB, each byte adds <= 255
    }
    if (p.smem_hist) flush();
    if (shorts) atomicAdd(gcnt + (size_t)(p.S + 1u) * kQualCounters, (unsigned long long)shorts);
}

// ================================================================================================
// 4. routing: stable partition of read ids by sample index
// ================================================================================================
struct RouteParams {
    const uint16_t *idx;
    uint64_t n;
    uint32_t bins;       // S + 1
    uint32_t ranges;     // one warp per range
    uint32_t chunk;      // reads per range (multiple of 32)
    uint32_t warps;      // per CTA
    uint32_t row_words;  // shared-memory words per warp: bins counters / cursors + bins lane slots (bytes)
    uint32_t vec;        // idx is 16-byte aligned and chunk a multiple of 8: the histogram loads eight indices at once
    uint32_t nbits;      // bits of a bin number, the idle lanes' `bins` included (ballot ranking)
    uint32_t *matrix;    // [ranges][bins]
    unsigned long long *totals;  // [bins], then the exclusive scan in place
    uint32_t *order;
    unsigned long long *offsets;  // [bins + 1]
};

__global__ void sgdx_route_hist(const RouteParams p) {
    extern __shared__ uint32_t s_route[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t range = blockIdx.x * p.warps + warp;
    if (range >= p.ranges) return;
    uint32_t *h = s_route + (size_t)warp * p.row_words;
    for (uint32_t b = lane; b < p.bins; b += 32u) h[b] = 0u;
    __syncwarp();
    const uint64_t lo = (uint64_t)range * p.chunk, hi = min(p.n, lo + p.chunk);
    const uint32_t last = p.bins - 1u;
    uint64_t r = lo;
    if (p.vec) {  // idx 16-byte aligned, chunk a multiple of 8: eight indices per lane and load (order is irrelevant here)
        const uint4 *src = reinterpret_cast<const uint4 *>(p.idx + lo);
        const uint64_t nvec = (hi - lo) >> 3;
#pragma unroll 2
        for (uint64_t i = lane; i < nvec; i += 32u) {
            const uint4 v = __ldg(src + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                atomicAdd(h + min(w[k] & 0xFFFFu, last), 1u);
                atomicAdd(h + min(w[k] >> 16, last), 1u);
            }
        }
        r = lo + (nvec << 3);
    }
#pragma unroll 4
    for (r += lane; r < hi; r += 32u) atomicAdd(h + min((uint32_t)__ldg(p.idx + r), last), 1u);
    __syncwarp();
    uint32_t *row = p.matrix + (size_t)range * p.bins;
    for (uint32_t b = lane; b < p.bins; b += 32u) row[b] = h[b];
}

// column b of the matrix becomes its exclusive prefix over the ranges; totals[b] = column sum.  One warp per bin,
// lanes stride over the ranges, shuffle scan with a running carry.
__global__ void __launch_bounds__(256) sgdx_route_scan_ranges(const RouteParams p) {
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (b >= p.bins) return;
    uint32_t carry = 0;
    for (uint32_t t0 = 0; t0 < p.ranges; t0 += 128u) {
        uint32_t c[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {  // independent strided loads first, then the four shuffle scans
            const uint32_t t = t0 + 32u * j + lane;
            c[j] = t < p.ranges ? p.matrix[(size_t)t * p.bins + b] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t t = t0 + 32u * j + lane;
            uint32_t incl = c[j];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl += up;
            }
            if (t < p.ranges) p.matrix[(size_t)t * p.bins + b] = carry + incl - c[j];
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    if (lane == 0) p.totals[b] = carry;
}

// one CTA: exclusive scan of the bin totals -> offsets[0..bins], totals[] := bin bases
__global__ void __launch_bounds__(1024) sgdx_route_scan_bins(const RouteParams p) {
    __shared__ unsigned long long s_part[1024];
    const uint32_t per = (p.bins + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(p.bins, lo + per);
    unsigned long long sum = 0;
    for (uint32_t b = lo; b < hi; b++) sum += p.totals[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (uint32_t i = 0; i < 1024u; i++) {
            const unsigned long long c = s_part[i];
            s_part[i] = run;
            run += c;
        }
    }
    __syncthreads();
    unsigned long long run = s_part[threadIdx.x];
    for (uint32_t b = lo; b < hi; b++) {
        const unsigned long long c = p.totals[b];
        p.totals[b] = run;
        p.offsets[b] = run;
        run += c;
    }
    if (threadIdx.x == 1023u) p.offsets[p.bins] = p.n;
}

// Ranked scatter.  A warp walks its range 32 reads at a time; lane i needs the lanes below it that hold the same bin.
// __match_any_sync answers that, but it executes on the address-divergence unit at about 64 cycles per instruction
// and SM, which made this kernel the slowest stage per byte.  Here the lanes of a bin find each other through a
// per-warp byte array in shared memory instead: every lane that has not been seen yet writes its lane number into
// slot[bin] (one of them wins), all lanes of the bin read the winner and add it to their peer mask; after as many
// rounds as the most frequent bin of the step has lanes (2-4 for a few hundred samples) every lane knows its peers.
// Steps dominated by one bin (few samples, or a run of Undetermined reads) stop after kRouteRounds rounds and take
// the hardware match; two to four bins are ranked by one vote per bin.
constexpr uint32_t kRouteRounds = 5;

__global__ void __launch_bounds__(256) sgdx_route_scatter(const RouteParams p) {
    extern __shared__ uint32_t s_route[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t range = blockIdx.x * p.warps + warp;
    if (range >= p.ranges) return;
    // cursors and lane slots of the bins, plus one dummy bin (index `bins`) for the lanes past the end of the range
    uint32_t *cur = s_route + (size_t)warp * p.row_words;
    uint8_t *slot = reinterpret_cast<uint8_t *>(cur + p.bins + 1u);
    const uint32_t *row = p.matrix + (size_t)range * p.bins;
    for (uint32_t b = lane; b < p.bins; b += 32u) cur[b] = row[b] + (uint32_t)p.totals[b];
    __syncwarp();
    const uint64_t lo = (uint64_t)range * p.chunk, hi = min(p.n, lo + p.chunk);
    const uint32_t lt = (1u << lane) - 1u, last = p.bins - 1u;
    const bool few = p.bins <= 4u;
    for (uint64_t base = lo; base < hi; base += 128u) {
        uint32_t bb[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {  // four independent loads in flight per lane
            const uint64_t r = base + 32u * j + lane;
            bb[j] = r < hi ? min((uint32_t)__ldg(p.idx + r), last) : p.bins;
        }
#pragma unroll 1
        for (int j = 0; j < 4; j++) {
            const uint32_t r = (uint32_t)base + 32u * j + lane;
            const uint32_t b = bb[j];
            uint32_t peers = 0u;
            if (few) {
                for (uint32_t k = 0; k <= p.bins; k++) {
                    const uint32_t v = __ballot_sync(0xFFFFFFFFu, b == k);
                    if (b == k) peers = v;
                }
            } else {
                uint32_t pending = 1u, rounds = 0u, more;
                do {
                    if (pending) slot[b] = (uint8_t)lane;
                    __syncwarp();
                    const uint32_t w = slot[b];
                    peers |= 1u << w;
                    pending = w != lane ? pending : 0u;
                    more = __any_sync(0xFFFFFFFFu, pending != 0u);  // (also orders this round's reads before the next writes)
                } while (more && ++rounds < kRouteRounds);
                if (more) peers = __match_any_sync(0xFFFFFFFFu, b);
            }
            const uint32_t rank = __popc(peers & lt);
            uint32_t at = 0;
            if (rank == 0u) {  // lowest lane of each bin group advances the bin's cursor
                at = cur[b];
                cur[b] = at + __popc(peers);
            }
            at = __shfl_sync(0xFFFFFFFFu, at, __ffs(peers) - 1);
            if (b != p.bins) p.order[at + rank] = r;
            __syncwarp();
        }
    }
}

// ---- CTA-tile form (the one that runs whenever its tables fit shared memory) ------------------------------------
// With one range per warp the scatter keeps ranges x bins output positions open at once; each sits in its own
// 128-byte line of L2, a few thousand warps times a few hundred bins overflow it, and half-written sectors go to
// DRAM and come back (measured at 50 M reads, 385 bins: 0.9 GB written for 0.2 GB of read ids).  Here a range belongs
// to a CTA of W warps that walks it in tiles of 32 W reads, warp w taking reads 32 w .. 32 w + 31 of the tile, so that
// the warps of a CTA write into the same lines and only (CTAs x bins) positions are open.
//   ranking inside a warp: every lane ORs its lane bit into a per-warp word of its bin (ATOMS.OR), reads the word
//     back -- the lanes of the bin -- and the bin's lowest lane clears it and publishes the count as byte
//     cnts[bin][warp];
//   after a barrier every lane reads the W count bytes of its bin: the bytes below its warp sum to its offset inside
//     the tile (IDP.4A against 0x01010101), all of them to the tile's total; position = cursor[bin] + offset + rank;
//   after a second barrier the lowest lane of the bin's last warp advances the cursor and clears the count bytes.
// The count bytes are double-buffered by tile parity, so two barriers per tile are enough.
constexpr uint64_t kRouteStagedFrom = 64ull << 20;  // reads per call from which the staged scatter is used (sgdx_route_device)

struct RouteTileParams {
    RouteParams r;
    uint32_t o_cnts, o_bm;  // byte offsets in shared memory (cursors at 0)
};

template <uint32_t W>
__host__ __device__ inline size_t route_tile_smem(uint32_t bins, uint32_t *o_cnts, uint32_t *o_bm) {
    size_t o = ((size_t)bins * 4u + 15u) & ~(size_t)15u;  // cursors / histogram
    if (o_cnts) *o_cnts = (uint32_t)o;
    o += (size_t)2u * bins * W;                            // count bytes, two tiles
    o = (o + 15u) & ~(size_t)15u;
    if (o_bm) *o_bm = (uint32_t)o;
    o += (size_t)W * (bins + 1u) * 4u;                     // lane words per warp, + the dummy bin
    return o;
}

// histogram of one CTA range (all warps into one shared histogram) -> matrix row
__global__ void __launch_bounds__(1024) sgdx_route_hist_cta(const RouteParams p) {
    extern __shared__ uint32_t s_route[];
    const uint32_t range = blockIdx.x;
    for (uint32_t b = threadIdx.x; b < p.bins; b += blockDim.x) s_route[b] = 0u;
    __syncthreads();
    const uint64_t lo = (uint64_t)range * p.chunk, hi = min(p.n, lo + p.chunk);
    const uint32_t last = p.bins - 1u;
    uint64_t r = lo;
    if (p.vec) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.idx + lo);
        const uint64_t nvec = (hi - lo) >> 3;
#pragma unroll 2
        for (uint64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
            const uint4 v = __ldg(src + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                atomicAdd(s_route + min(w[k] & 0xFFFFu, last), 1u);
                atomicAdd(s_route + min(w[k] >> 16, last), 1u);
            }
        }
        r = lo + (nvec << 3);
    }
    for (r += threadIdx.x; r < hi; r += blockDim.x) atomicAdd(s_route + min((uint32_t)__ldg(p.idx + r), last), 1u);
    __syncthreads();
    uint32_t *row = p.matrix + (size_t)range * p.bins;
    for (uint32_t b = threadIdx.x; b < p.bins; b += blockDim.x) row[b] = s_route[b];
}

template <uint32_t W>
__global__ void __launch_bounds__(W * 32u) sgdx_route_scatter_cta(const RouteTileParams q) {
    extern __shared__ __align__(16) unsigned char s_tile[];
    const RouteParams &p = q.r;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t range = blockIdx.x, bins = p.bins;
    uint32_t *cur = reinterpret_cast<uint32_t *>(s_tile);
    uint8_t *cnts = s_tile + q.o_cnts;
    uint32_t *bm = reinterpret_cast<uint32_t *>(s_tile + q.o_bm) + (size_t)warp * (bins + 1u);
    const uint32_t *row = p.matrix + (size_t)range * bins;
    for (uint32_t b = threadIdx.x; b < bins; b += W * 32u) cur[b] = row[b] + (uint32_t)p.totals[b];
    for (uint32_t i = threadIdx.x; i < (2u * bins * W) / 4u; i += W * 32u) reinterpret_cast<uint32_t *>(cnts)[i] = 0u;
    for (uint32_t b = lane; b <= bins; b += 32u) bm[b] = 0u;
    __syncthreads();
    const uint64_t lo = (uint64_t)range * p.chunk, hi = min(p.n, lo + p.chunk);
    const uint32_t lt = (1u << lane) - 1u, last = bins - 1u;
    // byte masks of the count words: bytes of the warps below mine
    uint32_t below[W / 4u];
#pragma unroll
    for (uint32_t k = 0; k < W / 4u; k++)
        below[k] = 4u * k + 4u <= warp ? 0xFFFFFFFFu : (4u * k >= warp ? 0u : (1u << (8u * (warp - 4u * k))) - 1u);
    // the index of the tile after the next one is requested at the top of an iteration and first looked at two
    // iterations later (a use right behind the load would stall the warp for a DRAM round trip in every tile)
    uint64_t r = lo + 32u * warp + lane;
    uint32_t b = r < hi ? min((uint32_t)__ldg(p.idx + r), last) : bins;
    uint32_t raw1 = r + 32u * W < hi ? (uint32_t)__ldg(p.idx + r + 32u * W) : 0xFFFFFFFFu;  // next tile, unclamped
    for (uint32_t t = 0; lo + (uint64_t)t * (32u * W) < hi; t++) {
        const uint64_t rn = r + 32u * W, rnn = r + 64u * W;
        const uint32_t raw2 = rnn < hi ? (uint32_t)__ldg(p.idx + rnn) : 0xFFFFFFFFu;
        uint8_t *cn = cnts + (size_t)(t & 1u) * bins * W;
        atomicOr(bm + b, 1u << lane);
        __syncwarp();
        const uint32_t peers = bm[b];
        const uint32_t rank = __popc(peers & lt), cnt = __popc(peers);
        const bool real = b != bins;
        __syncwarp();
        if (rank == 0u) {
            bm[b] = 0u;
            if (real) cn[b * W + warp] = (uint8_t)cnt;
        }
        __syncthreads();
        uint32_t prefix = 0u, tot = 0u;
        if (real) {
            const uint4 *cw = reinterpret_cast<const uint4 *>(cn + b * W);
            uint32_t x[W / 4u];
            if constexpr (W == 8u) {
                const uint2 v = *reinterpret_cast<const uint2 *>(cw);
                x[0] = v.x, x[1] = v.y;
            } else {
#pragma unroll
                for (uint32_t k = 0; k < W / 16u; k++) {
                    const uint4 v = cw[k];
                    x[4 * k] = v.x, x[4 * k + 1] = v.y, x[4 * k + 2] = v.z, x[4 * k + 3] = v.w;
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < W / 4u; k++) {
                prefix = __dp4a(x[k] & below[k], 0x01010101u, prefix);
                tot = __dp4a(x[k], 0x01010101u, tot);
            }
            p.order[cur[b] + prefix + rank] = (uint32_t)r;
        }
        __syncthreads();
        if (real && rank == 0u && prefix + cnt == tot) {  // the bin's last warp of this tile
            cur[b] += tot;
            if constexpr (W == 8u) {
                *reinterpret_cast<uint2 *>(cn + b * W) = make_uint2(0u, 0u);
            } else {
#pragma unroll
                for (uint32_t k = 0; k < W / 16u; k++) reinterpret_cast<uint4 *>(cn + b * W)[k] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        r = rn;
        b = raw1 == 0xFFFFFFFFu ? bins : min(raw1, last);
        raw1 = raw2;
    }
}

// ---- scatter, chained + staged form (large batches) -------------------------------------------------------------
// Beyond ~64 M reads the tile form above leaves its 8-9 us per million reads (1.44 ms at 100 M, 3.9 ms at 200 M): every
// id is its own 4-byte write into a 32-byte sector, the output spans more pages than the TLB reaches and more open
// lines than stay in L2, and the fragments travel.  This form sends whole sectors:
//   * ranking by votes: one ballot per bit of the bin number, the lanes that agree with me on every bit are my bin's
//     (no lane words in shared memory);
//   * no count matrix: the warps of a CTA visit the cursors in warp order.  Warp w waits for its turn on named
//     barrier w + 1 (`bar.sync id, 64`: its 32 threads + the 32 of the warp before it, which leaves with `bar.arrive
//     id, 64` -- the producer / consumer use of named barriers: shared-memory writes before the arrive are visible
//     after the sync), reads cursor[bin], the bin's lowest lane adds the warp's count, the turn goes on;
//   * the ids of a bin collect in a ring of eight shared-memory words per bin -- one sector of the output.  Positions
//     are counted from the sector boundary below `order` (P = position + a, a = words between that boundary and
//     order[0]), so sector s is P in [8s, 8s + 8).  A sector that a step fills from its first word goes straight to
//     global memory (eight lanes, one request); a sector that is not full after the step, or that an earlier step
//     began, goes to ring[bin][P & 7]; the bin's lowest lane sends the begun sector out (two STG.128) if the step
//     completes it.  Ring words are written and read inside the turn only.  The first sector of a range (shared
//     with the range before) and what is left in the rings at the end go out word by word.
// Measured (C2's 385 bins, ms for histogram + scans + scatter): 50 M reads 0.46 (tile form 0.44), 100 M 0.85 (1.44).
__device__ __forceinline__ void named_sync(uint32_t id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_arrive(uint32_t id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// rank of a lane among the lanes of its warp with the same bin, and their number
__device__ __forceinline__ void warp_rank_votes(uint32_t b, uint32_t lt, uint32_t nbits, uint32_t &rank, uint32_t &cnt) {
    uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t k = 0; k < 14u; k++) {  // (bins <= SGDX_MAX_SAMPLES + 1 < 2^14)
        if (k < nbits) {
            const bool bit = (b >> k) & 1u;
            const uint32_t v = __ballot_sync(0xFFFFFFFFu, bit);
            peers &= bit ? v : ~v;
        }
    }
    rank = __popc(peers & lt);
    cnt = __popc(peers);
}

constexpr uint32_t kStagedWarps = 8;  // (named barriers 1..8; barrier 0 stays __syncthreads')
inline size_t route_staged_smem(uint32_t bins) { return (size_t)bins * 32u + (size_t)2u * (bins + 1u) * 4u; }

__global__ void __launch_bounds__(kStagedWarps * 32u) sgdx_route_scatter_staged(const RouteParams p) {
    constexpr uint32_t W = kStagedWarps;
    extern __shared__ __align__(16) unsigned char s_tile[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t range = blockIdx.x, bins = p.bins;
    uint32_t *ring = reinterpret_cast<uint32_t *>(s_tile);                   // [bins][8]
    const uint4 *ring4 = reinterpret_cast<const uint4 *>(s_tile);
    uint32_t *cur = ring + (size_t)bins * 8u;                                // [bins + 1], the last one for idle lanes
    uint32_t *start = cur + (bins + 1u);                                     // [bins + 1]: the range's first P per bin
    const uint32_t a = (uint32_t)((reinterpret_cast<uintptr_t>(p.order) >> 2) & 7u);
    uint32_t *const order_a = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(p.order) - 4u * a);  // 32-byte aligned
    const uint32_t *row = p.matrix + (size_t)range * bins;
    for (uint32_t b = threadIdx.x; b <= bins; b += W * 32u) {
        const uint32_t c = b < bins ? row[b] + (uint32_t)p.totals[b] + a : 0u;
        cur[b] = c;
        start[b] = c;
    }
    __syncthreads();
    const uint64_t lo = (uint64_t)range * p.chunk, hi = min(p.n, lo + p.chunk);
    const uint32_t lt = (1u << lane) - 1u, last = bins - 1u;
    const uint32_t turn = warp + 1u, next = warp + 1u == W ? 1u : warp + 2u;
    if (warp == W - 1u) named_arrive(1u);  // warp 0 may start
    uint64_t r = lo + 32u * warp + lane;
    uint32_t b = r < hi ? min((uint32_t)__ldg(p.idx + r), last) : bins;
    uint32_t raw1 = r + 32u * W < hi ? (uint32_t)__ldg(p.idx + r + 32u * W) : 0xFFFFFFFFu;  // next tile, unclamped
    for (uint32_t t = 0; lo + (uint64_t)t * (32u * W) < hi; t++) {
        const uint64_t rn = r + 32u * W, rnn = r + 64u * W;
        const uint32_t raw2 = rnn < hi ? (uint32_t)__ldg(p.idx + rnn) : 0xFFFFFFFFu;
        uint32_t rank, cnt;
        warp_rank_votes(b, lt, p.nbits, rank, cnt);
        const bool real = b != bins;
        __syncwarp();  // (the named barriers are warp-aligned instructions)
        named_sync(turn);
        const uint32_t base = cur[b], end = base + cnt, P = base + rank, s8 = P & ~7u;
        if (rank == 0u && real) cur[b] = end;
        const bool complete = s8 + 8u <= end;  // the step fills my sector
        const bool begun = s8 < base;          // an earlier step began it (only the step's first sector can be)
        const bool to_ring = real && (!complete || begun);
        const bool first = s8 == (base & ~7u);
        uint32_t *const slot = ring + b * 8u + (P & 7u);
        if (to_ring && first) *slot = (uint32_t)r;
        __syncwarp();
        if (real && rank == 0u && begun && complete) {  // (P = base: my sector is the begun one)
            const uint32_t st = start[b];
            if (s8 >= st) {
                const uint4 v0 = ring4[b * 2u], v1 = ring4[b * 2u + 1u];
                uint4 *dst = reinterpret_cast<uint4 *>(order_a + s8);
                dst[0] = v0;
                dst[1] = v1;
            } else {  // the range's first sector of this bin: the words below `st` belong to the range before
                for (uint32_t j = st; j < s8 + 8u; j++) order_a[j] = ring[b * 8u + (j & 7u)];
            }
        }
        __syncwarp();
        if (to_ring && !first) *slot = (uint32_t)r;  // the sector the step leaves unfinished
        __syncwarp();
        named_arrive(next);
        if (real && !to_ring) order_a[P] = (uint32_t)r;
        r = rn;
        b = raw1 == 0xFFFFFFFFu ? bins : min(raw1, last);
        raw1 = raw2;
    }
    if (warp == 0u) named_sync(1u);  // the last warp's closing arrive
    __syncthreads();
    for (uint32_t bb = threadIdx.x; bb < bins; bb += W * 32u) {
        const uint32_t e = cur[bb];
        for (uint32_t j = max(e & ~7u, start[bb]); j < e; j++) order_a[j] = ring[bb * 8u + (j & 7u)];
    }
}

// ================================================================================================
// host side
// ================================================================================================
static UnmTable unm_table(const UnmatchedState &u) {
    UnmTable t;
    t.slots = u.d_slots;
    t.stats = u.d_stats;
    t.mask = u.cap_slots - 1;
    t.key_limit = u.key_limit;
    t.odd = u.d_odd;
    t.odd_cap = u.odd_cap;
    return t;
}

// exactly one of the three inputs is non-null: ASCII barcodes, compact records, 16-byte records
int post_after_match(sgdx_handle *h, Slot &s, const uint8_t *d_ascii, const uint64_t *d_compact, const uint4 *d_packed,
                     const uint16_t *d_idx, uint64_t n) {
    UnmatchedState &u = h->unm;
    const uint32_t grid = grid_cap(n, kUnmTile, h->sms, u.ctas_per_sm);  // exactly one wave of resident CTAs
    const uint32_t S = h->cfg.num_samples, L = h->cfg.barcode_len;
    if (d_ascii) sgdx_unmatched_count<0><<<grid, kUnmThreads, 0, s.stream>>>(d_ascii, d_idx, n, S, L, unm_table(u));
    else if (d_compact) sgdx_unmatched_count<1><<<grid, kUnmThreads, 0, s.stream>>>(d_compact, d_idx, n, S, L, unm_table(u));
    else sgdx_unmatched_count<2><<<grid, kUnmThreads, 0, s.stream>>>(d_packed, d_idx, n, S, L, unm_table(u));
    CU(cudaGetLastError());
    h->launches.fetch_add(1);
    if (!s.h_unm_stats) CU(cudaMallocHost(&s.h_unm_stats, kUnmStats * sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(s.h_unm_stats, u.d_stats, kUnmStats * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       s.stream));
    s.unm_pending = true;
    return SGDX_OK;
}

static void decode_key(ulonglong2 k, uint32_t L, char *out) {
    static const char letters[8] = {'A', 'C', 'T', 'G', '?', '?', '?', 'N'};
    const unsigned __int128 bits = ((unsigned __int128)(k.y & 0xFFFFFFFFull) << 64) | k.x;
    for (uint32_t j = 0; j < L; j++) out[j] = letters[(unsigned)(bits >> (3 * j)) & 7u];
}

// device idle, post_mu held: fold the foreign-byte list into the host map and empty it
static int unm_drain_odd(sgdx_handle *h) {
    UnmatchedState &u = h->unm;
    unsigned long long st[kUnmStats];
    CU(cudaMemcpy(st, u.d_stats, sizeof st, cudaMemcpyDeviceToHost));
    const uint64_t have = std::min<uint64_t>(st[3], u.odd_cap);
    if (have) {
        std::vector<uint8_t> buf(have * 32u);
        CU(cudaMemcpy(buf.data(), u.d_odd, buf.size(), cudaMemcpyDeviceToHost));
        const uint32_t L = h->cfg.barcode_len;
        for (uint64_t i = 0; i < have; i++) u.odd[std::string(reinterpret_cast<const char *>(&buf[i * 32u]), L)]++;
    }
    CU(cudaMemset(u.d_stats + 3, 0, sizeof(unsigned long long)));
    // the memset is on the default stream; the next batch runs on a non-blocking one and appends to this counter
    CU(cudaDeviceSynchronize());
    return SGDX_OK;
}

struct UnmEntry {
    std::string key;
    uint64_t count;
};

// device idle, post_mu held.  The `want` most frequent device-table entries (all of them when want is huge), found
// by a count threshold from the device histogram; only those are copied to the host.  keep_on_device: leave the
// compacted (key, count) list in d_sel_* and report its length.
static int unm_select_top(sgdx_handle *h, uint64_t want, std::vector<ulonglong2> &keys, std::vector<unsigned long long> &counts) {
    UnmatchedState &u = h->unm;
    unsigned long long st[kUnmStats];
    CU(cudaMemcpy(st, u.d_stats, sizeof st, cudaMemcpyDeviceToHost));
    const uint64_t distinct = st[0];
    keys.clear();
    counts.clear();
    if (distinct == 0 || want == 0) return SGDX_OK;
    unsigned long long t = 0, quota = 0;
    uint64_t need = distinct;
    if (want < distinct) {
        CU(cudaMemset(u.d_hist, 0, kUnmHistBins * sizeof(unsigned long long)));
        sgdx_unmatched_hist<<<grid_cap(u.cap_slots, 256, h->sms, 8), 256>>>(u.d_slots, u.cap_slots, u.d_hist);
        CU(cudaGetLastError());
        std::vector<unsigned long long> hist(kUnmHistBins);
        CU(cudaMemcpy(hist.data(), u.d_hist, hist.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        // largest t with #(count >= t) >= want; counts above the last bin are not resolved: take them all
        uint64_t above = 0;  // #(count > t) while walking down
        t = kUnmHistBins - 1;
        for (;;) {
            if (above + hist[t] >= want || t == 1) break;
            above += hist[t];
            t--;
        }
        if (t == (unsigned long long)(kUnmHistBins - 1)) {
            t = kUnmHistBins - 2;  // everything with count >= last bin
            quota = 0;
            need = hist[kUnmHistBins - 1];
        } else {
            quota = want > above ? want - above : 0;
            need = above + std::min<uint64_t>(quota, hist[t]);
        }
    }
    if (need > u.sel_cap) {
        cudaFree(u.d_sel_keys);
        cudaFree(u.d_sel_counts);
        u.d_sel_keys = nullptr;
        u.d_sel_counts = nullptr;
        u.sel_cap = 0;
        CU(cudaMalloc(&u.d_sel_keys, need * sizeof(ulonglong2)));
        CU(cudaMalloc(&u.d_sel_counts, need * sizeof(unsigned long long)));
        u.sel_cap = need;
    }
    CU(cudaMemset(u.d_stats + 4, 0, 2 * sizeof(unsigned long long)));
    sgdx_unmatched_select<<<grid_cap(u.cap_slots, 256, h->sms, 8), 256>>>(u.d_slots, u.cap_slots, t, quota,
                                                                          u.d_sel_keys, u.d_sel_counts, u.sel_cap,
                                                                          u.d_stats);
    CU(cudaGetLastError());
    CU(cudaMemcpy(st, u.d_stats, sizeof st, cudaMemcpyDeviceToHost));
    const uint64_t got = std::min<uint64_t>(st[4], u.sel_cap);
    keys.resize(got);
    counts.resize(got);
    if (got) {
        CU(cudaMemcpy(keys.data(), u.d_sel_keys, got * sizeof(ulonglong2), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(counts.data(), u.d_sel_counts, got * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    return SGDX_OK;
}

// the `want` most frequent keys of table + foreign-byte map, sorted by (count desc, key asc)
static int unm_top_entries(sgdx_handle *h, uint64_t want, std::vector<UnmEntry> &out) {
    if (int rc = unm_drain_odd(h)) return rc;
    std::vector<ulonglong2> keys;
    std::vector<unsigned long long> counts;
    if (int rc = unm_select_top(h, want, keys, counts)) return rc;
    const uint32_t L = h->cfg.barcode_len;
    out.clear();
    out.reserve(keys.size() + h->unm.odd.size());
    std::string k(L, 'A');
    for (size_t i = 0; i < keys.size(); i++) {
        decode_key(keys[i], L, &k[0]);
        out.push_back(UnmEntry{k, counts[i]});
    }
    for (auto &kv : h->unm.odd) out.push_back(UnmEntry{kv.first, kv.second});
    std::sort(out.begin(), out.end(), [](const UnmEntry &a, const UnmEntry &b) {
        return a.count != b.count ? a.count > b.count : a.key < b.key;
    });
    if (out.size() > want) out.resize(want);
    return SGDX_OK;
}

// UnmatchedCounter::downsize (metrics.rs:166-175); device idle, post_mu held
static int unm_downsize_locked(sgdx_handle *h) {
    UnmatchedState &u = h->unm;
    std::vector<UnmEntry> top;
    if (int rc = unm_top_entries(h, u.downsize_to, top)) return rc;
    std::vector<ulonglong2> keys;
    std::vector<unsigned long long> counts;
    std::map<std::string, uint64_t> odd;
    const uint32_t L = h->cfg.barcode_len;
    for (const UnmEntry &e : top) {
        unsigned __int128 bits = 0;
        bool foreign = false;
        for (uint32_t j = 0; j < L; j++) {
            const char c = e.key[j];
            const unsigned code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : c == 'G' ? 3 : c == 'N' ? 7 : 4;
            foreign |= code == 4;
            bits |= (unsigned __int128)code << (3 * j);
        }
        if (foreign) odd[e.key] = e.count;
        else {
            keys.push_back(make_ulonglong2((unsigned long long)bits, (1ull << 63) | (unsigned long long)(bits >> 64)));
            counts.push_back(e.count);
        }
    }
    u.odd.swap(odd);
    CU(cudaMemset(u.d_slots, 0xFF, u.cap_slots * sizeof(UnmSlot)));
    CU(cudaMemset(u.d_stats, 0, sizeof(unsigned long long)));  // distinct; the read totals stay
    if (!keys.empty()) {
        if (keys.size() > u.sel_cap) return fail(SGDX_ESTATE, "unmatched downsize: selection scratch too small");
        CU(cudaMemcpy(u.d_sel_keys, keys.data(), keys.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(u.d_sel_counts, counts.data(), counts.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        sgdx_unmatched_reinsert<<<grid_cap(keys.size(), 256, h->sms, 8), 256>>>(u.d_sel_keys, u.d_sel_counts, keys.size(),
                                                                               unm_table(u));
        CU(cudaGetLastError());
    }
    CU(cudaDeviceSynchronize());
    u.downsizes++;
    return SGDX_OK;
}

int post_on_sync(sgdx_handle *h, Slot &s) {
    if (!s.unm_pending) return SGDX_OK;
    s.unm_pending = false;
    UnmatchedState &u = h->unm;
    const unsigned long long *st = s.h_unm_stats;
    std::lock_guard<std::mutex> lk(h->post_mu);
    const bool too_many = st[0] + u.odd.size() >= u.max_map_size;
    const bool odd_filling = st[3] > u.odd_cap / 2;
    if (!too_many && !odd_filling) return SGDX_OK;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    if (too_many) return unm_downsize_locked(h);
    return unm_drain_odd(h);
}

int post_reset(sgdx_handle *h) {
    const uint32_t S = h->cfg.num_samples;
    if (h->qual.d_counters) CU(cudaMemset(h->qual.d_counters, 0, ((size_t)(S + 1) * kQualCounters + 1) * sizeof(unsigned long long)));
    UnmatchedState &u = h->unm;
    if (u.on) {
        CU(cudaMemset(u.d_slots, 0xFF, u.cap_slots * sizeof(UnmSlot)));
        CU(cudaMemset(u.d_stats, 0, kUnmStats * sizeof(unsigned long long)));
        u.odd.clear();
        u.downsizes = 0;
    }
    return SGDX_OK;
}

void post_destroy(sgdx_handle *h) {
    UnmatchedState &u = h->unm;
    cudaFree(u.d_slots);
    cudaFree(u.d_stats);
    cudaFree(u.d_odd);
    cudaFree(u.d_sel_keys);
    cudaFree(u.d_sel_counts);
    cudaFree(u.d_hist);
    cudaFree(h->qual.d_counters);
    for (Slot &s : h->slots) cudaFree(s.d_qmasks);
}

}  // namespace sgdx

// ------------------------------------------------------------------------------------------------
// C ABI: unmatched
// ------------------------------------------------------------------------------------------------
extern "C" int sgdx_unmatched_enable(sgdx_handle *h, uint64_t max_map_size, uint64_t downsize_to) {
    return sgdx::guarded("sgdx_unmatched_enable", [&]() -> int {
    if (!h) return fail(SGDX_EINVAL, "null handle");
    if (h->unm.on) return fail(SGDX_ESTATE, "unmatched collection is already enabled");
    if (h->cfg.barcode_len > SGDX_MAX_PACKED_LEN)
        return fail(SGDX_EINVAL, "the unmatched table keys barcodes of at most %u bases (96-bit keys)", SGDX_MAX_PACKED_LEN);
    if (max_map_size == 0) max_map_size = 5000000ull;  // opts.rs:185
    if (downsize_to == 0) downsize_to = 5000ull;       // opts.rs:189
    if (downsize_to > max_map_size) downsize_to = max_map_size;
    if (max_map_size > (1ull << 30)) return fail(SGDX_EINVAL, "max_map_size %llu too large", (unsigned long long)max_map_size);
    std::lock_guard<std::mutex> lk(h->post_mu);
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    UnmatchedState &u = h->unm;
    u.max_map_size = max_map_size;
    u.downsize_to = downsize_to;
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgdx_unmatched_count<1>, (int)kUnmThreads, 0));  // (the form with the most shared memory)
    u.ctas_per_sm = per_sm < 1 ? 1 : per_sm;
    // room for a batch's worth of new keys between two syncs: 4x the limit, at least 2^21 slots (64 MB)
    uint64_t cap = 1ull << 21;  // key_limit must stay above the publishing lag of the distinct-key count
    while (cap < 4 * max_map_size) cap <<= 1;
    u.cap_slots = cap;
    u.key_limit = cap / 4 * 3;
    u.odd_cap = 1ull << 20;
    CU(cudaMalloc(&u.d_slots, cap * sizeof(UnmSlot)));
    CU(cudaMalloc(&u.d_stats, kUnmStats * sizeof(unsigned long long)));
    CU(cudaMalloc(&u.d_odd, u.odd_cap * 32u));
    CU(cudaMalloc(&u.d_hist, kUnmHistBins * sizeof(unsigned long long)));
    CU(cudaMemset(u.d_slots, 0xFF, cap * sizeof(UnmSlot)));
    CU(cudaMemset(u.d_stats, 0, kUnmStats * sizeof(unsigned long long)));
    CU(cudaDeviceSynchronize());
    u.on = true;
    return SGDX_OK;
    });
}

#define UNM_ENTER(h)                                                                       \
    if (!(h)) return fail(SGDX_EINVAL, "null handle");                                     \
    if (!(h)->unm.on) return fail(SGDX_ESTATE, "call sgdx_unmatched_enable first");        \
    std::lock_guard<std::mutex> lk((h)->post_mu);                                          \
    CU(cudaSetDevice((h)->cfg.device));                                                    \
    CU(cudaDeviceSynchronize())

extern "C" int sgdx_unmatched_info_get(sgdx_handle *h, sgdx_unmatched_info *out) {
    return sgdx::guarded("sgdx_unmatched_info_get", [&]() -> int {
    if (!out) return fail(SGDX_EINVAL, "sgdx_unmatched_info_get: null out");
    UNM_ENTER(h);
    if (int rc = unm_drain_odd(h)) return rc;
    unsigned long long st[kUnmStats];
    CU(cudaMemcpy(st, h->unm.d_stats, sizeof st, cudaMemcpyDeviceToHost));
    out->distinct_keys = st[0] + h->unm.odd.size();
    out->unmatched_reads = st[1];
    out->dropped_reads = st[2];
    out->downsizes = h->unm.downsizes;
    return SGDX_OK;
    });
}

static int unm_write(sgdx_handle *h, const std::vector<UnmEntry> &v, uint8_t *keys, uint64_t *counts, uint64_t cap,
                     uint64_t *n_written) {
    if (v.size() > cap) return fail(SGDX_EINVAL, "unmatched export needs room for %zu keys, got %llu", v.size(), (unsigned long long)cap);
    const uint32_t L = h->cfg.barcode_len;
    for (size_t i = 0; i < v.size(); i++) {
        memcpy(keys + i * L, v[i].key.data(), L);
        counts[i] = v[i].count;
    }
    *n_written = v.size();
    return SGDX_OK;
}

extern "C" int sgdx_unmatched_export(sgdx_handle *h, uint8_t *keys, uint64_t *counts, uint64_t cap, uint64_t *n_written) {
    return sgdx::guarded("sgdx_unmatched_export", [&]() -> int {
    if (!n_written || (cap && (!keys || !counts))) return fail(SGDX_EINVAL, "sgdx_unmatched_export: null argument");
    UNM_ENTER(h);
    std::vector<UnmEntry> v;
    if (int rc = unm_top_entries(h, ~0ull, v)) return rc;
    return unm_write(h, v, keys, counts, cap, n_written);
    });
}

extern "C" int sgdx_unmatched_top(sgdx_handle *h, uint64_t n, uint8_t *keys, uint64_t *counts, uint64_t *n_written) {
    return sgdx::guarded("sgdx_unmatched_top", [&]() -> int {
    if (!n_written || (n && (!keys || !counts))) return fail(SGDX_EINVAL, "sgdx_unmatched_top: null argument");
    UNM_ENTER(h);
    std::vector<UnmEntry> v;
    if (int rc = unm_top_entries(h, n, v)) return rc;
    return unm_write(h, v, keys, counts, n, n_written);
    });
}

extern "C" int sgdx_unmatched_downsize(sgdx_handle *h) {
    return sgdx::guarded("sgdx_unmatched_downsize", [&]() -> int {
    UNM_ENTER(h);
    return unm_downsize_locked(h);
    });
}

// ------------------------------------------------------------------------------------------------
// C ABI: extraction
// ------------------------------------------------------------------------------------------------
static int fill_extract(sgdx_handle *h, uint32_t num_sources, const uint8_t *const *d_bases, const uint32_t *strides,
                        const sgdx_segment *segs, uint32_t nsegs, ExtractParams *p) {
    if (!d_bases || !strides || !segs) return fail(SGDX_EINVAL, "segment extraction: null argument");
    if (num_sources < 1 || num_sources > SGDX_MAX_SOURCES) return fail(SGDX_EINVAL, "1..%u sources", SGDX_MAX_SOURCES);
    if (nsegs < 1 || nsegs > SGDX_MAX_SEGMENTS) return fail(SGDX_EINVAL, "1..%u segments", SGDX_MAX_SEGMENTS);
    const uint32_t L = h->cfg.barcode_len;
    memset(p, 0, sizeof *p);
    for (uint32_t s = 0; s < num_sources; s++)
        if (!d_bases[s] || strides[s] == 0 || strides[s] > 0xFFFFu) return fail(SGDX_EINVAL, "source %u: null or stride outside 1..65535", s);
    // sample-barcode segments in FASTQ order, then segment order (demux.rs:504-522)
    uint32_t pos = 0;
    for (uint32_t s = 0; s < num_sources; s++)
        for (uint32_t i = 0; i < nsegs; i++) {
            const sgdx_segment &g = segs[i];
            if (g.source >= num_sources) return fail(SGDX_EINVAL, "segment %u: source %u of %u", i, g.source, num_sources);
            if (g.source != s || g.kind != SGDX_SEG_SAMPLE_BARCODE) continue;
            uint32_t len = g.length == SGDX_SEG_TO_END ? (g.offset < strides[s] ? strides[s] - g.offset : 0u) : g.length;
            if ((uint64_t)g.offset + len > strides[s]) return fail(SGDX_EINVAL, "segment %u runs past the stride of source %u", i, s);
            if (pos + len > L) return fail(SGDX_EINVAL, "sample-barcode segments are longer than barcode_len %u", L);
            if (len == 0) continue;
            p->run_ptr[p->runs] = d_bases[s] + g.offset;
            p->run_stride[p->runs] = strides[s];
            p->run_len[p->runs] = len;
            p->runs++;
            pos += len;
        }
    if (pos != L) return fail(SGDX_EINVAL, "sample-barcode segments sum to %u bases, barcode_len is %u (run.rs:109-119)", pos, L);
    p->L = L;
    return SGDX_OK;
}

static int check_slot_post(sgdx_handle *h, uint32_t slot) {
    if (!h) return fail(SGDX_EINVAL, "null handle");
    if (slot >= h->slots.size()) return fail(SGDX_EINVAL, "slot %u >= %zu", slot, h->slots.size());
    return SGDX_OK;
}

extern "C" int sgdx_extract_device(sgdx_handle *h, uint32_t slot, uint32_t num_sources, const uint8_t *const *d_bases,
                                   const uint32_t *strides, const sgdx_segment *segs, uint32_t nsegs, uint64_t n,
                                   uint8_t *d_out) {
    return sgdx::guarded("sgdx_extract_device", [&]() -> int {
    if (int rc = check_slot_post(h, slot)) return rc;
    ExtractParams p;
    if (int rc = fill_extract(h, num_sources, d_bases, strides, segs, nsegs, &p)) return rc;
    if (n == 0) return SGDX_OK;
    if (!d_out) return fail(SGDX_EINVAL, "sgdx_extract_device: null output");
    CU(cudaSetDevice(h->cfg.device));
    p.n = n;
    p.out = d_out;
    p.words_ok = (reinterpret_cast<uintptr_t>(d_out) & 3u) == 0u;
    sgdx_extract<0><<<grid_cap(n, kExtractThreads, h->sms, 8), kExtractThreads, 0, h->slots[slot].stream>>>(p);
    CU(cudaGetLastError());
    h->launches.fetch_add(1);
    return SGDX_OK;
    });
}

extern "C" int sgdx_match_segments_device(sgdx_handle *h, uint32_t slot, uint32_t num_sources,
                                          const uint8_t *const *d_bases, const uint32_t *strides,
                                          const sgdx_segment *segs, uint32_t nsegs, uint64_t n, uint16_t *d_idx,
                                          uint8_t *d_dist) {
    return sgdx::guarded("sgdx_match_segments_device", [&]() -> int {
    if (int rc = check_slot_post(h, slot)) return rc;
    Slot &s = h->slots[slot];
    CU(cudaSetDevice(h->cfg.device));
    const uint32_t L = h->cfg.barcode_len;
    if (n > s.extract_cap) {
        CU(cudaStreamSynchronize(s.stream));
        cudaFree(s.d_extract);
        s.d_extract = nullptr;
        s.extract_cap = 0;
        const uint64_t cap = n + n / 8 + 1024;
        CU(cudaMalloc(&s.d_extract, cap * std::max<uint32_t>(L, 16u)));  // L ASCII bytes, or an 8- / 16-byte record
        s.extract_cap = cap;
    }
    // Handles whose tables run the perfect-hash kernels take the barcodes as records, written by the extraction
    // kernel itself.  With the unmatched table on the ASCII form is kept: a barcode with a byte that is none of
    // A/C/G/T/N has a key only there.
    const bool compact = L <= kCompactMaxLen && (h->single || (h->ph.slots && !h->ph.wide));
    const bool packed = !compact && L <= SGDX_MAX_PACKED_LEN && h->ph.slots && h->ph.wide;
    if ((compact || packed) && !h->unm.on) {
        ExtractParams p;
        if (int rc = fill_extract(h, num_sources, d_bases, strides, segs, nsegs, &p)) return rc;
        if (n) {
            p.n = n;
            p.out = s.d_extract;
            if (compact) sgdx_extract<1><<<grid_cap(n, kExtractThreads, h->sms, 8), kExtractThreads, 0, s.stream>>>(p);
            else sgdx_extract<2><<<grid_cap(n, kExtractThreads, h->sms, 8), kExtractThreads, 0, s.stream>>>(p);
            CU(cudaGetLastError());
            h->launches.fetch_add(1);
        }
        return compact ? sgdx_match_compact_device(h, slot, reinterpret_cast<const uint64_t *>(s.d_extract), n, d_idx, d_dist)
                       : sgdx_match_packed_device(h, slot, reinterpret_cast<const sgdx_read *>(s.d_extract), n, d_idx, d_dist);
    }
    if (int rc = sgdx_extract_device(h, slot, num_sources, d_bases, strides, segs, nsegs, n, s.d_extract)) return rc;
    return sgdx_match_device(h, slot, s.d_extract, n, d_idx, d_dist);
    });
}

// ------------------------------------------------------------------------------------------------
// C ABI: base-quality counters
// ------------------------------------------------------------------------------------------------
static int ensure_qual_counters(sgdx_handle *h) {
    if (h->qual.d_counters) return SGDX_OK;
    std::lock_guard<std::mutex> lk(h->post_mu);
    if (h->qual.d_counters) return SGDX_OK;
    const size_t words = (size_t)(h->cfg.num_samples + 1) * kQualCounters + 1;
    unsigned long long *d = nullptr;
    CU(cudaMalloc(&d, words * sizeof(unsigned long long)));
    CU(cudaMemset(d, 0, words * sizeof(unsigned long long)));
    CU(cudaDeviceSynchronize());
    h->qual.d_counters = d;
    return SGDX_OK;
}

extern "C" int sgdx_qual_counts_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_quals, uint64_t n,
                                       uint32_t stride, const uint32_t *d_lens, const sgdx_segment *segs,
                                       uint32_t nsegs, const uint16_t *d_idx) {
    return sgdx::guarded("sgdx_qual_counts_device", [&]() -> int {
    if (int rc = check_slot_post(h, slot)) return rc;
    if (!segs || nsegs < 1 || nsegs > SGDX_MAX_SEGMENTS) return fail(SGDX_EINVAL, "sgdx_qual_counts_device: 1..%u segments", SGDX_MAX_SEGMENTS);
    if (stride < 1 || stride > 3072u) return fail(SGDX_EINVAL, "stride %u outside 1..3072", stride);
    if (n && (!d_quals || !d_idx)) return fail(SGDX_EINVAL, "sgdx_qual_counts_device: null buffer");
    if (reinterpret_cast<uintptr_t>(d_quals) & 15u) return fail(SGDX_EINVAL, "quality matrix must be 16-byte aligned");
    CU(cudaSetDevice(h->cfg.device));
    if (int rc = ensure_qual_counters(h)) return rc;
    Slot &s = h->slots[slot];

    // position masks of the structure: bit p of tbits / bbits = position p is a template / sample-barcode base
    const uint32_t nwords = (stride + 15u + 31u) / 32u + 1u;
    std::vector<uint32_t> key;
    key.push_back(stride);
    uint32_t min_len = 0, has_b = 0;
    std::vector<uint8_t> kind(stride, 0);
    for (uint32_t i = 0; i < nsegs; i++) {
        const sgdx_segment &g = segs[i];
        key.insert(key.end(), {g.offset, g.length, g.kind});
        if (g.kind != SGDX_SEG_TEMPLATE && g.kind != SGDX_SEG_SAMPLE_BARCODE && g.kind != SGDX_SEG_MOLECULAR_BARCODE &&
            g.kind != SGDX_SEG_SKIP)
            return fail(SGDX_EINVAL, "segment %u: unknown kind %u", i, g.kind);
        uint32_t end;
        if (g.length == SGDX_SEG_TO_END) {
            if (g.offset > stride) return fail(SGDX_EINVAL, "segment %u starts past the stride", i);
            end = stride;
            min_len = std::max(min_len, g.offset);
        } else {
            if ((uint64_t)g.offset + g.length > stride) return fail(SGDX_EINVAL, "segment %u runs past the stride", i);
            end = g.offset + g.length;
            min_len = std::max(min_len, end);
        }
        for (uint32_t q = g.offset; q < end; q++) {
            if (kind[q]) return fail(SGDX_EINVAL, "segments overlap at position %u", q);
            kind[q] = (uint8_t)g.kind;
        }
        if (g.kind == SGDX_SEG_SAMPLE_BARCODE && end > g.offset) has_b = 1;
    }
    if (key != s.qmasks_key) {
        std::vector<uint32_t> masks((size_t)32 * nwords, 0u);
        for (uint32_t a = 0; a < 16; a++)
            for (uint32_t q = 0; q < stride; q++) {
                const uint32_t bit = q + a;
                if (kind[q] == SGDX_SEG_TEMPLATE) masks[(size_t)a * nwords + (bit >> 5)] |= 1u << (bit & 31u);
                if (kind[q] == SGDX_SEG_SAMPLE_BARCODE) masks[(size_t)(16 + a) * nwords + (bit >> 5)] |= 1u << (bit & 31u);
            }
        if (masks.size() > s.qmasks_cap) {
            CU(cudaStreamSynchronize(s.stream));  // a kernel still reading the old masks
            cudaFree(s.d_qmasks);
            s.d_qmasks = nullptr;
            s.qmasks_cap = 0;
            s.qmasks_key.clear();
            CU(cudaMalloc(&s.d_qmasks, masks.size() * sizeof(uint32_t)));
            s.qmasks_cap = masks.size();
        }
        // on the slot's stream: ordered after the kernels that read the old masks and before the one queued
        // below (a copy on the default stream orders against neither); pageable source, staged before return
        CU(cudaMemcpyAsync(s.d_qmasks, masks.data(), masks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
        s.qmasks_key = key;
    }
    if (n == 0) return SGDX_OK;

    QualParams p{};
    p.quals = d_quals;
    p.lens = d_lens;
    p.idx = d_idx;
    p.n = n;
    p.stride = stride;
    p.S = h->cfg.num_samples;
    p.nwords = nwords;
    p.min_len = min_len;
    p.one = 1u;
    p.short_rows = stride <= 32u ? 1u : 0u;
    if (p.short_rows)
        for (uint32_t q = 0; q < stride; q++) {
            if (kind[q] == SGDX_SEG_TEMPLATE) p.tmask32 |= 1u << q;
            if (kind[q] == SGDX_SEG_SAMPLE_BARCODE) p.bmask32 |= 1u << q;
        }
    p.t_tail = stride;
    while (p.t_tail > 0 && kind[p.t_tail - 1] == SGDX_SEG_TEMPLATE) p.t_tail--;
    p.masks = s.d_qmasks;
    p.counters = h->qual.d_counters;
    // tile = reads_per_tile quality strings, a multiple of 16 reads so that every tile starts 16-byte aligned
    const uint32_t threads = p.short_rows ? kQualShortThreads : kQualThreads, lanes = threads / kQualSplit;
    uint32_t R = lanes;
    while (R > 16u && (uint64_t)R * stride > 48u * 1024u) R >>= 1;
    if ((uint64_t)R * stride < 8u * 1024u) R = lanes * std::max<uint32_t>(1u, std::min<uint32_t>(32u, 16u * 1024u / (lanes * stride)));  // short reads: ~16 KB tiles
    p.reads_per_tile = R;
    p.stage_bytes = (R * stride + 16u + 127u) & ~127u;
    p.stages = 2;
    const size_t hist_bytes = (size_t)(p.S + 1) * kQualCounters * sizeof(uint32_t);
    p.smem_hist = hist_bytes <= 48u * 1024u ? 1u : 0u;
    const size_t nbins = p.smem_hist ? (size_t)(p.S + 1) * kQualCounters : 0;
    const size_t smem = (size_t)p.stages * p.stage_bytes + (size_t)32 * nwords * 4 + ((nbins + 1) & ~(size_t)1) * 4 + p.stages * 8;
    auto kern = p.short_rows ? (has_b ? sgdx_qual_counts<true, true> : sgdx_qual_counts<false, true>)
                             : (has_b ? sgdx_qual_counts<true, false> : sgdx_qual_counts<false, false>);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (int)threads, smem));
    if (per_sm < 1) return fail(SGDX_EINVAL, "quality-counter kernel does not fit (stride %u, %u samples)", stride, p.S);
    const uint32_t grid = grid_cap(n, R, h->sms, per_sm);
    kern<<<grid, threads, smem, s.stream>>>(p);
    CU(cudaGetLastError());
    h->launches.fetch_add(1);
    return SGDX_OK;
    });
}

extern "C" int sgdx_base_qual_counters(sgdx_handle *h, sgdx_base_qual *per_sample, uint64_t *short_reads) {
    return sgdx::guarded("sgdx_base_qual_counters", [&]() -> int {
    if (!h || !per_sample) return fail(SGDX_EINVAL, "sgdx_base_qual_counters: null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    const uint32_t S = h->cfg.num_samples;
    std::vector<unsigned long long> raw((size_t)(S + 1) * kQualCounters + 1, 0ull);
    if (h->qual.d_counters) CU(cudaMemcpy(raw.data(), h->qual.d_counters, raw.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (uint32_t b = 0; b <= S; b++) {
        per_sample[b].q20_bases = raw[(size_t)b * kQualCounters + 0];
        per_sample[b].q30_bases = raw[(size_t)b * kQualCounters + 1];
        per_sample[b].index_bases_qual_sum = raw[(size_t)b * kQualCounters + 2];
        per_sample[b].index_bases_total_seen = raw[(size_t)b * kQualCounters + 3];
        per_sample[b].bases = raw[(size_t)b * kQualCounters + 4];
    }
    if (short_reads) *short_reads = raw.back();
    return SGDX_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// C ABI: routing
// ------------------------------------------------------------------------------------------------
extern "C" int sgdx_route_device(sgdx_handle *h, uint32_t slot, const uint16_t *d_idx, uint64_t n, uint32_t *d_order,
                                 uint64_t *d_offsets) {
    return sgdx::guarded("sgdx_route_device", [&]() -> int {
    if (int rc = check_slot_post(h, slot)) return rc;
    if (!d_offsets || (n && (!d_idx || !d_order))) return fail(SGDX_EINVAL, "sgdx_route_device: null buffer");
    if (n >> 32) return fail(SGDX_EINVAL, "sgdx_route_device: %llu reads, at most 2^32 - 1 per call", (unsigned long long)n);
    CU(cudaSetDevice(h->cfg.device));
    Slot &s = h->slots[slot];
    RouteParams p{};
    p.idx = d_idx;
    p.n = n;
    p.bins = h->cfg.num_samples + 1;
    const size_t row_bytes = (size_t)p.bins * sizeof(uint32_t);
    p.vec = (reinterpret_cast<uintptr_t>(d_idx) & 15u) == 0u ? 1u : 0u;  // (chunk is a multiple of 32 reads = 64 bytes)
    // CTA-tile form when its tables fit shared memory (two CTAs per SM if possible), else per-warp ranges (W = 0).
    // Measured at 50 M reads, 385 bins (ms for histogram + scans + scatter): W = 32: 0.58, W = 16: 0.44, W = 8: 0.48
    // -- the count bytes of wide CTAs cost shared-memory wavefronts, narrow ones need more CTAs for the same warps.
    uint32_t W = 0;
    uint64_t per_sm = 12u;
    if (route_tile_smem<16>(p.bins, nullptr, nullptr) <= 113u * 1024u) W = 16, per_sm = 2u;
    else if (route_tile_smem<8>(p.bins, nullptr, nullptr) <= 113u * 1024u) W = 8, per_sm = 2u;
    // (more than ~2200 bins: one 8-warp CTA per SM would be all that fits -- the per-warp ranges have more warps)
    // large batches: the chained + staged form, four CTAs per SM (measured: 4 beats 7) -- when its rings fit
    bool staged = false;
    p.nbits = 32u - (uint32_t)__builtin_clz(p.bins);
    {
        uint64_t staged_from = kRouteStagedFrom;
#ifndef SGDX_NO_TEST_HOOKS  // test / measurement hook: the smallest batch that takes the staged form
        if (const char *e = getenv("SGDX_ROUTE_STAGED_FROM")) staged_from = strtoull(e, nullptr, 10);
#endif
        const size_t cta_smem = route_staged_smem(p.bins) + 1024u;
        const uint32_t ctas = (uint32_t)std::min<size_t>(4u, (227u * 1024u) / cta_smem);
        if (n >= staged_from && ctas >= 2u && n + 8u < (1ull << 32)) staged = true, W = kStagedWarps, per_sm = ctas;
    }
#ifndef SGDX_NO_TEST_HOOKS  // measurement hook: ranges (= CTAs of the tile forms) per SM
    if (const char *e = getenv("SGDX_ROUTE_PER_SM"))
        if (atoi(e) > 0) per_sm = (uint64_t)atoi(e);
#endif
    if (W) {
        p.warps = W;
    } else {
        // one warp per contiguous range of reads; each warp keeps a private row of `bins` counters (cursors) and
        // `bins` one-byte lane slots in shared memory
        p.row_words = (p.bins + 1u) + (p.bins + 4u) / 4u;  // + the dummy bin of sgdx_route_scatter
        p.warps = (uint32_t)std::max<size_t>(1, std::min<size_t>(8, (96u * 1024u) / ((size_t)p.row_words * sizeof(uint32_t))));
    }
    // few enough ranges that the (ranges x bins) offset matrix stays small next to the index array itself
    const uint64_t max_ranges = (uint64_t)h->sms * per_sm;
    const uint64_t min_chunk = W ? 32u * W : 1024u;
    uint64_t ranges = std::min<uint64_t>(max_ranges, (n + min_chunk - 1) / min_chunk);
    if (ranges == 0) ranges = 1;
    uint64_t chunk = (n + ranges - 1) / ranges;
    chunk = (chunk + 31) & ~31ull;
    if (chunk == 0) chunk = 32;
    ranges = n ? (n + chunk - 1) / chunk : 1;
    p.ranges = (uint32_t)ranges;
    p.chunk = (uint32_t)chunk;
    const size_t need = ranges * row_bytes + (size_t)p.bins * sizeof(unsigned long long);
    if (need > s.route_cap) {
        CU(cudaStreamSynchronize(s.stream));
        cudaFree(s.d_route);
        s.d_route = nullptr;
        s.route_cap = 0;
        CU(cudaMalloc(&s.d_route, need + need / 4));
        s.route_cap = need + need / 4;
    }
    p.totals = static_cast<unsigned long long *>(s.d_route);  // 8-byte aligned part first
    p.matrix = reinterpret_cast<uint32_t *>(p.totals + p.bins);
    p.order = d_order;
    p.offsets = reinterpret_cast<unsigned long long *>(d_offsets);
    if (W) {
        RouteTileParams q{};
        q.r = p;
        const size_t smem = W == 16 ? route_tile_smem<16>(p.bins, &q.o_cnts, &q.o_bm) : route_tile_smem<8>(p.bins, &q.o_cnts, &q.o_bm);
        CU(cudaFuncSetAttribute(sgdx_route_hist_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes));
        sgdx_route_hist_cta<<<p.ranges, 1024, row_bytes, s.stream>>>(p);
        CU(cudaGetLastError());
        sgdx_route_scan_ranges<<<(p.bins + 7u) / 8u, 256, 0, s.stream>>>(p);
        CU(cudaGetLastError());
        sgdx_route_scan_bins<<<1, 1024, 0, s.stream>>>(p);
        CU(cudaGetLastError());
        if (n && staged) {
            const size_t csmem = route_staged_smem(p.bins);
            CU(cudaFuncSetAttribute(sgdx_route_scatter_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
            sgdx_route_scatter_staged<<<p.ranges, kStagedWarps * 32u, csmem, s.stream>>>(p);
            CU(cudaGetLastError());
        } else if (n) {
            auto kern = W == 16 ? sgdx_route_scatter_cta<16> : sgdx_route_scatter_cta<8>;
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<p.ranges, W * 32u, smem, s.stream>>>(q);
            CU(cudaGetLastError());
        }
        h->launches.fetch_add(n ? 4 : 3);
        return SGDX_OK;
    }
    const size_t smem = (size_t)p.warps * p.row_words * sizeof(uint32_t);
    CU(cudaFuncSetAttribute(sgdx_route_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaFuncSetAttribute(sgdx_route_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t grid = (p.ranges + p.warps - 1) / p.warps;
    sgdx_route_hist<<<grid, p.warps * 32u, smem, s.stream>>>(p);
    CU(cudaGetLastError());
    sgdx_route_scan_ranges<<<(p.bins + 7u) / 8u, 256, 0, s.stream>>>(p);
    CU(cudaGetLastError());
    sgdx_route_scan_bins<<<1, 1024, 0, s.stream>>>(p);
    CU(cudaGetLastError());
    if (n) {
        sgdx_route_scatter<<<grid, p.warps * 32u, smem, s.stream>>>(p);
        CU(cudaGetLastError());
    }
    h->launches.fetch_add(n ? 4 : 3);
    return SGDX_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// C ABI: one call per batch
// ------------------------------------------------------------------------------------------------
extern "C" int sgdx_demultiplex_device(sgdx_handle *h, uint32_t slot, const sgdx_fastq_view *fastqs, uint32_t num_fastqs,
                                       const sgdx_segment *segs, uint32_t nsegs, uint64_t n, const sgdx_demux_out *out) {
    return sgdx::guarded("sgdx_demultiplex_device", [&]() -> int {
    if (int rc = check_slot_post(h, slot)) return rc;
    if (!fastqs || !segs || !out) return fail(SGDX_EINVAL, "sgdx_demultiplex_device: null argument");
    if (num_fastqs < 1 || num_fastqs > SGDX_MAX_SOURCES) return fail(SGDX_EINVAL, "1..%u FASTQs", SGDX_MAX_SOURCES);
    if (nsegs < 1 || nsegs > SGDX_MAX_SEGMENTS) return fail(SGDX_EINVAL, "1..%u segments", SGDX_MAX_SEGMENTS);
    if (n && !out->d_sample_idx) return fail(SGDX_EINVAL, "sgdx_demultiplex_device: null sample index output");
    if (out->d_order && !out->d_offsets) return fail(SGDX_EINVAL, "sgdx_demultiplex_device: d_order needs d_offsets");
    const uint8_t *bases[SGDX_MAX_SOURCES];
    uint32_t strides[SGDX_MAX_SOURCES];
    for (uint32_t f = 0; f < num_fastqs; f++) {
        bases[f] = fastqs[f].d_bases;
        strides[f] = fastqs[f].stride;
    }
    // demux.rs:504-522 (barcode of the template) + 524-556 (assignment, counters, hop check, unmatched)
    if (int rc = sgdx_match_segments_device(h, slot, num_fastqs, bases, strides, segs, nsegs, n, out->d_sample_idx,
                                            out->d_hamming_dist))
        return rc;
    // demux.rs:360-417 (per-segment quality counters) + 582-583 (added to the template's sample)
    for (uint32_t f = 0; f < num_fastqs; f++) {
        if (!fastqs[f].d_quals) continue;
        sgdx_segment own[SGDX_MAX_SEGMENTS];
        uint32_t k = 0;
        for (uint32_t i = 0; i < nsegs; i++)
            if (segs[i].source == f) {
                own[k] = segs[i];
                own[k++].source = 0;
            }
        if (k == 0) continue;
        if (int rc = sgdx_qual_counts_device(h, slot, fastqs[f].d_quals, n, fastqs[f].stride, fastqs[f].d_lens, own, k,
                                             out->d_sample_idx))
            return rc;
    }
    // demux.rs:578: per_sample_reads[sample_idx].push(read), as a permutation
    if (out->d_order) return sgdx_route_device(h, slot, out->d_sample_idx, n, out->d_order, out->d_offsets);
    return SGDX_OK;
    });
}
