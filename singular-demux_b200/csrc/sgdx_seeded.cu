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
#include "sgdx_exact_first.cuh"

namespace sgdx {

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

size_t exact_smem_bytes(const DevParams &p, uint32_t threads) { return exact_smem_bytes_impl(p, threads); }
uint32_t exact_block_threads(const DevParams &p) { return exact_block_threads_impl(p); }

#define SGDX_DECL_WORDS(n) cudaError_t launch_exact_words_##n(const DevParams &, uint32_t, int, cudaStream_t);
SGDX_DECL_WORDS(1) SGDX_DECL_WORDS(2) SGDX_DECL_WORDS(3) SGDX_DECL_WORDS(4)
SGDX_DECL_WORDS(5) SGDX_DECL_WORDS(6) SGDX_DECL_WORDS(7) SGDX_DECL_WORDS(8)

cudaError_t launch_seeded(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    if (!packed && p.ex_on) {
        switch (p.Lw) {
            case 1: return launch_exact_words_1(p, mode, sms, stream);
            case 2: return launch_exact_words_2(p, mode, sms, stream);
            case 3: return launch_exact_words_3(p, mode, sms, stream);
            case 4: return launch_exact_words_4(p, mode, sms, stream);
            case 5: return launch_exact_words_5(p, mode, sms, stream);
            case 6: return launch_exact_words_6(p, mode, sms, stream);
            case 7: return launch_exact_words_7(p, mode, sms, stream);
            case 8: return launch_exact_words_8(p, mode, sms, stream);
            default: return cudaErrorInvalidValue;
        }
    }
    if (mode == kModeHamming)
        return packed ? launch_seeded_t<kModeHamming, true>(p, sms, stream)
                      : launch_seeded_t<kModeHamming, false>(p, sms, stream);
    return packed ? launch_seeded_t<kModePreCompute, true>(p, sms, stream)
                  : launch_seeded_t<kModePreCompute, false>(p, sms, stream);
}

}  // namespace sgdx
