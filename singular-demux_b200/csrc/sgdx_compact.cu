// sgdx_compact.cu -- assignment on compact records: one 64-bit word per read (barcodes of at most 21 bases).
//
//   bits  0..20  lo plane   bit j = bit 0 of base j's 2-bit code, code = (ascii >> 1) & 3  (A0 C1 T2 G3)
//   bits 21..41  hi plane   bit j = bit 1 of the code
//   bits 42..62  invalid    bit j = base j is not A/C/G/T; at such a position lo = 0 says 'N' (a no-call,
//                           demux.rs:78-80), lo = 1 says any other byte (mismatches every sample); hi = 0
//   bit  63      zero
//
// This is BASELINE.json's "2-bit base planes plus an N mask" and SURVEY 8(d)'s algorithmic input
// (ceil(3L/8) bytes per read): the host writes it instead of the L ASCII bytes when it fills the pinned batch
// buffer (sgdx_pack_compact_host, csrc/sgdx_hostpack.cpp), so a C2 read crosses PCIe as 8 bytes, not 20, and
// the kernel fetches it with one fully coalesced 64-bit load.  Same results as every other assignment kernel
// (matcher.rs:273-348 / 121-262, demux.rs:524-538, metrics.rs:428-437, 658-676).
//
// sgdx_match_compact has the shape of sgdx_match_exact_first (sgdx_exact_first.cuh) without its ASCII work:
//   front end (all reads)  a read without invalid positions is looked up in a cuckoo table of the expected
//                          barcodes (two positions from two multiply-xor hashes of the record's words, 64-bit
//                          entries = record | sample << 42); an equal entry is Match(sample, 0) (valid because
//                          min pairwise table distance > min_delta) -> done; everything else -> Q1
//   stage 2 (Q1)           reads without invalid positions: pigeonhole seed search -> done; others -> Q3
//   stage 3 (Q3)           seed search with 4-way expansion of an invalid base; windows with several -> one warp
//                          per read over all samples
// Every stage runs with the whole CTA on up to blockDim queued reads, one thread per read.
#include "sgdx_seed_search.cuh"

namespace sgdx {

constexpr uint32_t kCompactThreads = 512;
constexpr uint32_t kCompactReadsPerThread = 4;

struct CompactSmem {
    uint2 *table;     // [S] barcode planes
    uint32_t *q1;     // [(R+1)*threads] block-local read numbers
    uint32_t *q3;     // [2*threads]
    uint2 *slots;     // [1 << bits] 64-bit cuckoo entries
    uint32_t *hist;   // [(S+1)*3]
    uint32_t *cq;     // [threads] Q3 entries that need the all-samples scan
    uint32_t *ring;   // [3] pushes of the current phase
    uint32_t *sdesc;  // [8] chunk descriptors of the short seed index (a run-time index into the kernel
                      //     parameter would put the whole struct on the local stack)
    uint16_t *hdr;    // seed index
    uint16_t *ids;
    uint16_t *shdr;   // short seed index (stage 2), may be empty
    uint16_t *sids;
};

__host__ __device__ inline size_t compact_carve(unsigned char *base, uint32_t threads, uint32_t S, uint32_t bits,
                                                uint32_t hdr_total, uint32_t ids_total, uint32_t shdr_total,
                                                uint32_t sids_total, CompactSmem *out) {
    size_t o = 0;
    CompactSmem s;
    s.table = reinterpret_cast<uint2 *>(base + o);      o += ((size_t)S * 8 + 15) & ~(size_t)15;
    s.slots = reinterpret_cast<uint2 *>(base + o);      o += (size_t)8 << bits;
    s.q1 = reinterpret_cast<uint32_t *>(base + o);      o += (size_t)threads * (kCompactReadsPerThread + 1) * 4;
    s.q3 = reinterpret_cast<uint32_t *>(base + o);      o += (size_t)threads * 2 * 4;
    s.hist = reinterpret_cast<uint32_t *>(base + o);    o += (((size_t)(S + 1) * 3 * 4) + 15) & ~(size_t)15;
    s.cq = reinterpret_cast<uint32_t *>(base + o);      o += (size_t)threads * 4;
    s.ring = reinterpret_cast<uint32_t *>(base + o);    o += 16;
    s.sdesc = reinterpret_cast<uint32_t *>(base + o);   o += 32;
    s.hdr = reinterpret_cast<uint16_t *>(base + o);     o += ((size_t)hdr_total * 2 + 15) & ~(size_t)15;
    s.ids = reinterpret_cast<uint16_t *>(base + o);     o += ((size_t)ids_total * 2 + 15) & ~(size_t)15;
    s.shdr = reinterpret_cast<uint16_t *>(base + o);    o += ((size_t)shdr_total * 2 + 15) & ~(size_t)15;
    s.sids = reinterpret_cast<uint16_t *>(base + o);    o += ((size_t)sids_total * 2 + 15) & ~(size_t)15;
    if (out) *out = s;
    return o;
}

size_t compact_smem_bytes(const DevParams &p, uint32_t bits, uint32_t shdr_total, uint32_t sids_total) {
    return compact_carve(nullptr, kCompactThreads, p.S, bits, p.seed_hdr_total, p.seed_ids_total, shdr_total, sids_total,
                         nullptr);
}

template <uint32_t MODE>
__global__ void __launch_bounds__(kCompactThreads, 3) sgdx_match_compact(const DevParams p, const CompactParams c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t T = kCompactThreads, R = kCompactReadsPerThread, nwarps = T / 32;
    constexpr uint32_t kNone = 0xFFFFFFFFu;
    CompactSmem sm;
    compact_carve(smem_raw, T, p.S, c.bits, p.seed_hdr_total, p.seed_ids_total, c.short_hdr_total, c.short_ids_total, &sm);
    const uint32_t nbins = (p.S + 1u) * 3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t nslots = 1u << c.bits;

    for (uint32_t i = threadIdx.x; i < nslots; i += T) sm.slots[i] = reinterpret_cast<const uint2 *>(c.slots)[i];
    for (uint32_t i = threadIdx.x; i < nbins; i += T) sm.hist[i] = 0u;
    for (uint32_t i = threadIdx.x; i < p.S; i += T) sm.table[i] = p.table[i];
    for (uint32_t i = threadIdx.x; i < p.seed_hdr_total; i += T) sm.hdr[i] = p.seed_hdr[i];
    for (uint32_t i = threadIdx.x; i < p.seed_ids_total; i += T) sm.ids[i] = p.seed_ids[i];
    for (uint32_t i = threadIdx.x; i < c.short_hdr_total; i += T) sm.shdr[i] = c.short_hdr[i];
    for (uint32_t i = threadIdx.x; i < c.short_ids_total; i += T) sm.sids[i] = c.short_ids[i];
    if (threadIdx.x < 4) sm.ring[threadIdx.x] = 0u;
#pragma unroll
    for (uint32_t i = 0; i < 8; i++)
        if (threadIdx.x == i) sm.sdesc[i] = c.short_desc[i];
    __syncthreads();

    const uint32_t shift = 32u - c.bits;
    const uint32_t lmask = (1u << p.L) - 1u;  // barcode_len <= 21; stray bits above it are ignored
    uint32_t n1 = 0, n3 = 0;  // queue fills, tracked identically by every thread

    // queue entries number the block's reads: iteration * (T*R) + offset inside the iteration's tile
    auto read_of = [&](uint32_t e) -> uint64_t {
        return ((uint64_t)blockIdx.x + (uint64_t)(e / (T * R)) * gridDim.x) * (T * R) + (e % (T * R));
    };
    // one barrier per phase: it publishes the phase's pushes, everybody reads their number after it, and thread 0
    // clears the counter phase+2 will use (same protocol as sgdx_match_exact_first)
    uint32_t phase = 0;
    auto push_slot = [&]() -> uint32_t { return atomicAdd(&sm.ring[phase % 3u], 1u); };
    auto end_phase = [&]() -> uint32_t {
        __syncthreads();
        const uint32_t pushed = sm.ring[phase % 3u];
        if (threadIdx.x == 0) sm.ring[(phase + 2u) % 3u] = 0u;
        phase++;
        return pushed;
    };

    // ---- stage 3: reads with invalid positions, and whatever the seed index cannot enumerate
    auto stage3 = [&](uint32_t count) {
        const bool valid = threadIdx.x < count;
        uint64_t r = 0;
        Rd rd = {0u, 0u, 0u, 0u};
        uint32_t best = kInfKey, nxt = kInfKey, e = 0;
        bool scan = false;
        if (valid) {
            e = sm.q3[n3 - count + threadIdx.x];
            r = read_of(e);
            rd = rd_of_compact(__ldg(c.reads + r), lmask);
            if (!cannot_match<MODE>(p, rd))
                scan = !seed_search<MODE, true>(p, sm.hdr, sm.ids, sm.table, rd, best, nxt);
        }
        if (scan) sm.cq[push_slot()] = e;
        Verdict v = apply_rules<MODE>(p, rd, best, nxt);
        emit(p, r, valid && !scan, rd, v, sm.hist);
        const uint32_t nscan = end_phase();
        n3 -= count;
        if (nscan) {
            for (uint32_t q = warp; q < nscan; q += nwarps) {  // one warp per read, all samples
                const uint64_t qrid = read_of(sm.cq[q]);
                const Rd qrd = rd_of_compact(__ldg(c.reads + qrid), lmask);  // same address in every lane
                Verdict qv = warp_scan<MODE>(p, sm.table, qrd, lane);
                if (lane == 0) emit_single(p, qrid, qrd, qv, sm.hist);
            }
            __syncthreads();  // cq is free again before a following stage 3 pushes into it
        }
    };

    // ---- stage 2: seed search for reads made of A/C/G/T only
    auto stage2 = [&](uint32_t count) {
        const bool valid = threadIdx.x < count;
        uint64_t r = 0;
        Rd rd = {0u, 0u, 0u, 0u};
        uint32_t best = kInfKey, nxt = kInfKey;
        bool done = false;
        if (valid) {
            const uint32_t e = sm.q1[n1 - count + threadIdx.x];
            r = read_of(e);
            const uint64_t rec = __ldg(c.reads + r);
            if (((rec >> (2u * kCompactField)) & lmask) != 0ull) {
                sm.q3[n3 + push_slot()] = e;
            } else if (c.short_C) {
                // dmin > 2M + delta: at most one sample is within M of this read, and then every other one is
                // further than M + delta away -- the first candidate with d <= M is the verdict, nxt stays "none"
                rd = rd_of_compact(rec, lmask);
                for (uint32_t ch = 0; ch < c.short_C; ch++) {
                    const uint32_t desc = sm.sdesc[ch];
                    const uint32_t pos = desc & 0xFFu, w = (desc >> 8) & 0xFFu, off = desc >> 16;
                    const uint32_t wm = (1u << w) - 1u;
                    const uint32_t key = ((rd.lo >> pos) & wm) | (((rd.hi >> pos) & wm) << w);
                    const uint32_t start = sm.shdr[off + key], end = sm.shdr[off + key + 1u];
#pragma unroll 1
                    for (uint32_t i = start; i < end; i++) {
                        const uint32_t s = sm.sids[i];
                        const uint2 t = sm.table[s];
                        const uint32_t d = __popc((rd.lo ^ t.x) | (rd.hi ^ t.y));
                        if (d <= p.M) best = (d << 16) | s;
                    }
                }
                done = true;
            } else {  // an all-A/C/G/T read passes the gate and the cannot-match test
                rd = rd_of_compact(rec, lmask);
                seed_search<MODE, false>(p, sm.hdr, sm.ids, sm.table, rd, best, nxt);
                done = true;
            }
        }
        Verdict v = apply_rules<MODE>(p, rd, best, nxt);
        emit(p, r, done, rd, v, sm.hist);
        n3 += end_phase();
        n1 -= count;
    };

    uint32_t iter = 0;
    uint64_t base = (uint64_t)blockIdx.x * (T * R);
    for (;;) {
        const bool more = base < p.n;
        if (n3 >= T || (!more && n1 == 0u && n3)) {
            stage3(min(n3, T));
        } else if (n1 >= T || (!more && n1)) {
            stage2(min(n1, T));
        } else if (more) {
            uint64_t rec[R];
            uint32_t hit[R];
            const uint32_t nvalid = (uint32_t)min((uint64_t)(T * R), p.n - base);
#pragma unroll
            for (uint32_t k = 0; k < R; k++)  // lane i of a warp reads record base + k*T + 32*warp + i: 256 B per request
                rec[k] = k * T + threadIdx.x < nvalid ? __ldg(c.reads + base + k * T + threadIdx.x) : ~0ull;
#pragma unroll
            for (uint32_t k = 0; k < R; k++) {
                const uint64_t r = base + k * T + threadIdx.x;
                const uint32_t w0 = (uint32_t)rec[k], w1 = (uint32_t)(rec[k] >> 32);
                hit[k] = kNone;
                if ((w1 >> (2u * kCompactField - 32u)) == 0u) {  // no invalid position (and not a padding lane)
                    // two multiply-xor hashes of the record's words pick the two cuckoo positions; an entry is
                    // record | (sample + 1) << 42, so one 64-bit compare both verifies the barcode and yields
                    // the sample (an empty slot is 0: its sample field fails the test even for an all-A read)
                    const uint32_t a = ((w0 * c.k0) ^ (w1 * c.k1)) >> shift, b = ((w0 * c.k2) ^ (w1 * c.k3)) >> shift;
                    const uint2 e1 = sm.slots[a], e2 = sm.slots[b];
                    const uint32_t x1 = e1.y ^ w1, x2 = e2.y ^ w1;
                    const bool h1 = ((e1.x ^ w0) | (x1 & 0x3FFu)) == 0u && x1 != 0u;
                    const bool h2 = ((e2.x ^ w0) | (x2 & 0x3FFu)) == 0u && x2 != 0u;
                    if (h1 || h2) hit[k] = ((h1 ? x1 : x2) >> 10) - 1u;
                }
                if (hit[k] != kNone) {
                    p.out_idx[r] = (uint16_t)hit[k];
                    if (p.out_dist) p.out_dist[r] = 0u;
                    atomicAdd(&sm.hist[hit[k] * 3u + 1u], 1u);  // perfect match
                }
            }
            {   // misses -> Q1: one reservation per warp, the lanes take their slots from the ballots
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
        } else {
            break;
        }
    }

    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nbins; i += T) {
        uint32_t cnt = sm.hist[i];
        if (cnt) atomicAdd(&p.counters[i], (unsigned long long)cnt);
    }
}

// ASCII -> compact records (device-resident pipelines and tests; the host packer writes the same words)
__global__ void __launch_bounds__(256) sgdx_pack_compact(const uint8_t *__restrict__ reads, uint64_t n, uint32_t L,
                                                         uint64_t *__restrict__ out) {
    const bool words_ok = (L & 3u) == 0u && (reinterpret_cast<uintptr_t>(reads) & 3u) == 0u;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x)
        out[r] = compact_of_rd(load_ascii(reads, r, L, words_ok));
}

// compact -> 16-byte sgdx_read records, for tables the fast kernel does not cover (no seed index, or
// min pairwise distance <= min_delta): those go through the packed form of the general kernels
__global__ void __launch_bounds__(256) sgdx_expand_compact(const uint64_t *__restrict__ in, uint64_t n, uint32_t lmask,
                                                           uint4 *__restrict__ out) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const Rd rd = rd_of_compact(__ldg(in + r), lmask);
        out[r] = make_uint4(rd.lo, rd.hi, rd.nmask, rd.xmask);
    }
}

cudaError_t launch_compact(const DevParams &p, const CompactParams &c, uint32_t mode, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const size_t smem = compact_smem_bytes(p, c.bits, c.short_hdr_total, c.short_ids_total);
    auto kern = mode == kModeHamming ? sgdx_match_compact<kModeHamming> : sgdx_match_compact<kModePreCompute>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (int)kCompactThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    kern<<<persistent_grid(p.n, kCompactThreads * kCompactReadsPerThread, sms, per_sm), kCompactThreads, smem, stream>>>(p, c);
    return cudaGetLastError();
}

cudaError_t launch_pack_compact(const uint8_t *reads, uint64_t n, uint32_t L, uint64_t *out, int sms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    sgdx_pack_compact<<<persistent_grid(n, 256, sms, 8), 256, 0, stream>>>(reads, n, L, out);
    return cudaGetLastError();
}

cudaError_t launch_expand_compact(const uint64_t *in, uint64_t n, uint32_t L, uint4 *out, int sms, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    sgdx_expand_compact<<<persistent_grid(n, 256, sms, 8), 256, 0, stream>>>(in, n, (1u << L) - 1u, out);
    return cudaGetLastError();
}

}  // namespace sgdx
