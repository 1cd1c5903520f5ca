// sgdx_seeded.cu -- the seed-and-verify assignment kernel (sm_100a).
//
// Same results as sgdx_match_direct (matcher.rs:273-348 / 121-262, demux.rs:524-538, metrics.rs:428-437,
// 658-676), but instead of scoring every (read, sample) pair it enumerates only the samples that can
// influence the verdict.  A sample matters only if its number of mismatches over the read's valid
// (non-N, non-foreign) positions is <= T, with T = max_mismatches + min_delta for the Hamming rule set
// and T = max(M, M+delta-1) for the PreCompute rule set ("why T is enough").  Cut the
// barcode into T+1 disjoint chunks: by pigeonhole such a sample agrees with the read on every valid
// position of at least one chunk, so it sits in the bucket that chunk's bases select.  Buckets come from
// a small index built once per matcher (sgdx_create) and staged in shared memory next to the table.
//
//   - chunk window without invalid positions      -> one bucket
//   - chunk window with exactly one invalid base  -> the 4 buckets of its possible values
//   - anything worse (rare)                       -> the read is queued and scored against all samples
//                                                    by a whole warp (per-lane best/second-best keys,
//                                                    __reduce_min_sync across the warp)
// Reads that cannot match whatever the table holds (too many no-calls / foreign bytes for M, or the
// max_no_calls gate) skip the search.  Candidates are scored exactly like the direct kernel:
// popc(((lo^slo)|inv)|(hi^shi)), key = distance << 16 | sample; a sample reached through several
// chunks produces the same key twice, which the `key != best` guard makes harmless.
#include "sgdx_kernels.h"

namespace sgdx {

struct SeedSmem {
    uint4 *dq_rd;        // [threads] planes of the reads queued for the all-samples scan
    uint2 *table;        // [S]
    uint32_t *hdr;       // [hdr_total]
    uint32_t *hist;      // [(S+1)*3]
    uint32_t *dq_tid;    // [threads]
    uint32_t *dq_count;  // [1] (+3 pad)
    uint16_t *ids;       // [ids_total]
};

__host__ __device__ inline size_t seed_carve(unsigned char *base, uint32_t threads, uint32_t S, uint32_t hdr_total,
                                             uint32_t ids_total, SeedSmem *out) {
    size_t o = 0;
    SeedSmem s;
    s.dq_rd = reinterpret_cast<uint4 *>(base + o);      o += (size_t)threads * 16;
    s.table = reinterpret_cast<uint2 *>(base + o);      o += (size_t)S * 8;
    s.hdr = reinterpret_cast<uint32_t *>(base + o);     o += (size_t)hdr_total * 4;
    s.hist = reinterpret_cast<uint32_t *>(base + o);    o += (size_t)(S + 1) * 3 * 4;
    s.dq_tid = reinterpret_cast<uint32_t *>(base + o);  o += (size_t)threads * 4;
    s.dq_count = reinterpret_cast<uint32_t *>(base + o); o += 16;
    s.ids = reinterpret_cast<uint16_t *>(base + o);     o += ((size_t)ids_total * 2 + 15) & ~(size_t)15;
    if (out) *out = s;
    return o;
}

size_t seeded_smem_bytes(const DevParams &p) {
    return seed_carve(nullptr, kSeededThreads, p.S, p.seed_hdr_total, p.seed_ids_total, nullptr);
}

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

// verdict of one read written by one thread (the warp-cooperative path)
__device__ __forceinline__ void emit_single(const DevParams &p, uint64_t r, const Rd &rd, const Verdict &v,
                                            uint32_t *hist) {
    p.out_idx[r] = (uint16_t)v.bin;
    if (p.out_dist) p.out_dist[r] = (uint8_t)v.dist;
    uint32_t cls = v.dist == 0u ? 1u : (v.dist == 1u ? 2u : 0u);
    atomicAdd(&hist[v.bin * 3u + cls], 1u);
    if (p.hop_on && v.bin == p.S) hop_check(p, rd);
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
    const uint32_t C = p.seed_C;

    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = base + threadIdx.x;
        bool active = r < p.n;
        Rd rd = {0u, 0u, 0u, 0u};
        if (active) rd = PACKED ? load_packed(p.reads_packed, r, p.L) : load_ascii(p.reads_ascii, r, p.L, words_ok);
        const uint32_t inv = rd.nmask | rd.xmask;
        const uint32_t n_nc = __popc(rd.nmask), n_x = __popc(rd.xmask);
        const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;

        // reads no table can match: every sample is further than M away / the no-call gate (demux.rs:528-531)
        bool dead = p.max_no_calls >= 0 && n_nc > (uint32_t)p.max_no_calls;
        if (MODE == kModeHamming) dead |= n_x + (n_nc > p.free_ns ? n_nc - p.free_ns : 0u) > p.M;
        else dead |= rd.xmask != 0u || n_nc > p.M;

        uint32_t best = kInfKey, nxt = kInfKey;
        bool queued = false;
        if (active && !dead) {
            for (uint32_t c = 0; c < C; c++) {
                const uint32_t desc = p.seed_desc[c];
                const uint32_t pos = desc & 0xFFu, w = (desc >> 8) & 0xFFu, off = desc >> 16;
                const uint32_t wm = (1u << w) - 1u;
                const uint32_t iv = (inv >> pos) & wm;
                const uint32_t key0 = ((rd.lo >> pos) & wm) | (((rd.hi >> pos) & wm) << w);
                uint32_t nvar = 1u, j = 0u;
                if (iv) {
                    if (iv & (iv - 1u)) {  // two or more invalid bases in one window: leave it to a warp
                        queued = true;
                        break;
                    }
                    nvar = 4u;
                    j = __ffs(iv) - 1u;
                }
                for (uint32_t b = 0; b < nvar; b++) {
                    const uint32_t key = key0 | ((b & 1u) << j) | ((b >> 1) << (w + j));
                    const uint32_t h = sm.hdr[off + key];
                    const uint32_t start = h >> 14, cnt = h & 0x3FFFu;
                    for (uint32_t i = 0; i < cnt; i++) {
                        const uint32_t s = sm.ids[start + i];
                        const uint2 t = sm.table[s];
                        score<MODE>(rd, inv, skip, t.x, t.y, s, best, nxt);
                    }
                }
            }
        }
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
                const uint32_t qinv = qrd.nmask | qrd.xmask;
                const uint32_t qskip = (MODE == kModePreCompute && qrd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
                uint32_t b = kInfKey, n2 = kInfKey;
                for (uint32_t s = lane; s < p.S; s += 32u) {
                    const uint2 t = sm.table[s];
                    score<MODE>(qrd, qinv, qskip, t.x, t.y, s, b, n2);
                }
                const uint32_t gb = __reduce_min_sync(0xFFFFFFFFu, b);
                const uint32_t gn = __reduce_min_sync(0xFFFFFFFFu, b == gb ? n2 : b);  // keys are unique per sample
                if (lane == 0) {
                    Verdict qvd = apply_rules<MODE>(p, qrd, gb, gn);
                    emit_single(p, base + sm.dq_tid[q], qrd, qvd, sm.hist);
                }
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

static uint32_t seeded_grid(uint64_t n, int sms, int per_sm) {
    uint64_t tiles = (n + kSeededThreads - 1) / kSeededThreads;
    uint64_t cap = (uint64_t)sms * (uint64_t)per_sm;  // persistent grid: a multiple of the SM count
    return (uint32_t)(tiles < cap ? (tiles ? tiles : 1) : cap);
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
    kern<<<seeded_grid(p.n, sms, per_sm), kSeededThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_seeded(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    if (mode == kModeHamming)
        return packed ? launch_seeded_t<kModeHamming, true>(p, sms, stream)
                      : launch_seeded_t<kModeHamming, false>(p, sms, stream);
    return packed ? launch_seeded_t<kModePreCompute, true>(p, sms, stream)
                  : launch_seeded_t<kModePreCompute, false>(p, sms, stream);
}

}  // namespace sgdx
