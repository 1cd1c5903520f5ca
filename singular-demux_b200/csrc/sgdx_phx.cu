// sgdx_phx.cu -- sgdx_match_phx: the general form of sgdx_match_ph (sgdx_ph.cu) -- assignment on packed records through
// a perfect hash of the match neighbourhood (sm_100a).  sgdx_match_ph is the instantiation the headline configuration
// runs (compact records, key radius == max_mismatches), kept as its own lean kernel; this file serves 16-byte records
// (barcodes of 22..32 bases) and tables whose full neighbourhood does not fit (radius < max_mismatches).
//
// Same results as every other assignment kernel (matcher.rs:273-348 / 121-262, demux.rs:524-538,
// metrics.rs:428-437, 658-676).  Input: one 64-bit compact record per read (barcodes of at most 21 bases), or one
// 16-byte sgdx_read record per read (WIDE, up to 32 bases) -- both "2-bit base planes + invalid masks" (include/sgdx.h).
//
// Idea.  For a read made of A/C/G/T only, Matcher::find is a pure function of the string.  sgdx_create enumerates the
// strings within `radius` substitutions of an expected barcode (radius = max_mismatches when that set is small enough
// -- 23 424 strings for 384 samples of 20 bases at one mismatch -- else as large as fits, down to 0 = the barcodes
// themselves), decides each one with a full scan under the handle's rule set (best / second best, ties, delta,
// PreCompute's visibility rules: exactly what the general kernels do per read) and stores "sample or none" in a
// two-level perfect hash (PhParams, sgdx_kernels.h).  Per read the kernel does
//     three multiplies -> bucket multiplier (LDS.U16) -> entry (LDS.U16) -> that sample's planes (LDS.64)
//     -> XOR / OR / POPC -> distance <= radius ?
// The distance test is both the verification of the untagged entry and the reported hamming_dist.  A read that
// passes is decided.  One that fails is NoMatch when radius = max_mismatches (it is further than that from every
// sample) and undecided otherwise.  There is no search, no queue and no barrier on this path; everything else -- the
// index-hop check of NoMatch reads, reads with invalid positions, undecided reads -- is deferred to per-warp queues
// in shared memory and drained 32 at a time by the full warp (hop key sets and seed index in shared memory too, when
// they fit).
//
// Layout.  One 1024-thread CTA per SM, warps are independent (no __syncthreads between the table load and the final
// counter flush).  A warp takes chunks of consecutive reads: lane i owns 4 (WIDE: 2) consecutive reads, fetched by one
// 256-bit load and answered by one 64-bit + one 32-bit (WIDE: 32-bit + 16-bit) store.  The next chunk's records are
// requested before the current chunk is processed.  Buffers that are not aligned for this, and the last partial
// chunk, take the same code with narrower accesses (!VEC: lane i owns reads i, i+32, ...).
#include "sgdx_seed_search.cuh"

namespace sgdx {
namespace phx {

constexpr uint32_t kPhThreads = 1024;
constexpr uint32_t kPhWarps = kPhThreads / 32;
constexpr uint32_t kPhQueue2 = 64;  // <= 31 left over + <= 32 pushed by one drain of the first queue
constexpr size_t kPhSmemLimit = 227u * 1024u;

struct PhSmem {
    uint64_t *qrec;   // [warps][qcap] deferred reads: planes (compact record, or lo | hi << 32) ...
    uint32_t *qoff;   // [warps][qcap] ... and (warp iteration * chunk + offset in chunk)
    uint64_t *q2rec;  // [warps][kPhQueue2] with a short radius: the plain reads the hash levels left undecided ...
    uint32_t *q2off;  // [warps][kPhQueue2] ... same numbering
    uint64_t *q3rec;  // [warps][kPhQueue2] the reads with invalid positions or stray bits
    uint32_t *q3off;
    uint64_t *hslots; // hop key sets (PhParams::hop_slots), optional
    uint2 *planes;    // [S + 1][copies]: barcode planes, replicated so that the lanes of a half warp read distinct banks;
                      //                  row S = dummy
    uint32_t *hist;   // [3][S + 1] perfect, one mismatch, other (metrics.rs:428-437)
    uint16_t *disp;   // [1 << bbits]
    uint16_t *slots;  // [1 << sbits]
    uint16_t *hdisp;  // hop key sets
    uint16_t *hdr;    // seed index of DevParams (general path of undecided reads), optional
    uint16_t *ids;
    uint16_t *mhdr;   // short seed index of PhParams (ms_*), optional
    uint16_t *mids;
};

struct PhFit {       // shared-memory layout of one launch (ph_fit)
    uint32_t qcap;   // entries of a warp's first queue: <= 31 left over + the reads of one chunk
    uint32_t q2;     // 1 = the second queue exists (radius < max_mismatches)
    uint32_t clog;   // log2 of the copies of the planes table
    uint32_t hsbits, hbbits;  // hop key sets: slot bits (0 = not in shared memory), bucket bits
    uint32_t hdr_total, ids_total;  // seed index entries (0 = not in shared memory)
    uint32_t mhdr_total, mids_total;  // short seed index entries (0 = not in shared memory)
    // byte offsets of the arrays of PhSmem, filled on the host (ph_layout) so that the kernel adds instead of carving
    uint32_t o_hist, o_disp, o_slots, o_qrec, o_q2rec, o_q3rec, o_hslots, o_qoff, o_q2off, o_q3off, o_hdisp, o_hdr, o_ids, o_mhdr, o_mids;
};

__host__ __device__ inline size_t ph_carve(unsigned char *base, uint32_t S, uint32_t bbits, uint32_t sbits, const PhFit &f,
                                           PhSmem *out) {
    size_t o = 0;
    PhSmem s;
    // what the fast path touches first, at offsets that are cheap to form: the planes table at 0
    s.planes = reinterpret_cast<uint2 *>(base + o);    o += (((size_t)S + 1) * 8) << f.clog;
    s.hist = reinterpret_cast<uint32_t *>(base + o);   o += (((size_t)S + 1) * 3 * 4 + 15) & ~(size_t)15;
    s.disp = reinterpret_cast<uint16_t *>(base + o);   o += (size_t)2 << bbits;
    s.slots = reinterpret_cast<uint16_t *>(base + o);  o += (size_t)2 << sbits;
    o = (o + 15) & ~(size_t)15;
    s.qrec = reinterpret_cast<uint64_t *>(base + o);   o += (size_t)kPhWarps * f.qcap * 8;
    s.q2rec = reinterpret_cast<uint64_t *>(base + o);  o += f.q2 ? (size_t)kPhWarps * kPhQueue2 * 8 : 0;
    s.q3rec = reinterpret_cast<uint64_t *>(base + o);  o += (size_t)kPhWarps * kPhQueue2 * 8;
    s.hslots = reinterpret_cast<uint64_t *>(base + o); o += f.hsbits ? (size_t)8 << f.hsbits : 0;
    s.qoff = reinterpret_cast<uint32_t *>(base + o);   o += (size_t)kPhWarps * f.qcap * 4;
    s.q2off = reinterpret_cast<uint32_t *>(base + o);  o += f.q2 ? (size_t)kPhWarps * kPhQueue2 * 4 : 0;
    s.q3off = reinterpret_cast<uint32_t *>(base + o);  o += (size_t)kPhWarps * kPhQueue2 * 4;
    s.hdisp = reinterpret_cast<uint16_t *>(base + o);  o += f.hsbits ? (size_t)2 << f.hbbits : 0;
    s.hdr = reinterpret_cast<uint16_t *>(base + o);    o += ((size_t)f.hdr_total * 2 + 15) & ~(size_t)15;
    s.ids = reinterpret_cast<uint16_t *>(base + o);    o += ((size_t)f.ids_total * 2 + 15) & ~(size_t)15;
    s.mhdr = reinterpret_cast<uint16_t *>(base + o);   o += ((size_t)f.mhdr_total * 2 + 15) & ~(size_t)15;
    s.mids = reinterpret_cast<uint16_t *>(base + o);   o += ((size_t)f.mids_total * 2 + 15) & ~(size_t)15;
    o = (o + 15) & ~(size_t)15;
    if (out) *out = s;
    return o;
}

// the offsets of a layout, for the kernel
static void ph_layout(uint32_t S, uint32_t bbits, uint32_t sbits, PhFit &f) {
    PhSmem s;
    ph_carve(nullptr, S, bbits, sbits, f, &s);
    auto at = [](const void *q) { return (uint32_t)reinterpret_cast<uintptr_t>(q); };
    f.o_hist = at(s.hist), f.o_disp = at(s.disp), f.o_slots = at(s.slots), f.o_qrec = at(s.qrec), f.o_q2rec = at(s.q2rec);
    f.o_q3rec = at(s.q3rec), f.o_hslots = at(s.hslots), f.o_qoff = at(s.qoff), f.o_q2off = at(s.q2off), f.o_q3off = at(s.q3off);
    f.o_hdisp = at(s.hdisp), f.o_hdr = at(s.hdr), f.o_ids = at(s.ids), f.o_mhdr = at(s.mhdr), f.o_mids = at(s.mids);
}

// What is optional goes into shared memory as far as it fits, in the order of what it saves: the short seed index
// (every read the hash levels leave undecided searches it), the hop key sets (without them every hop check is a round
// trip to L2), the full seed index (without a short one every undecided read searches it; with one, only the few reads
// the short one cannot decide), copies of the planes table (up to one per lane of a half warp).
static PhFit ph_fit(const DevParams &p, const PhParams &c) {
    const uint32_t bbits = 32u - c.bshift, sbits = 32u - c.sshift;
    PhFit f{};
    f.qcap = c.wide ? 96u : 160u;
    f.q2 = c.radius < p.M ? 1u : 0u;
    f.hbbits = 32u - c.hop_bshift;
    auto fits = [&](const PhFit &x) { return ph_carve(nullptr, p.S, bbits, sbits, x, nullptr) <= kPhSmemLimit; };
    if (c.radius < p.M && c.ms_C) {
        PhFit g = f;
        g.mhdr_total = c.ms_hdr_total;
        g.mids_total = c.ms_ids_total;
        if (fits(g)) f = g;
    }
    const bool full_first = c.radius < p.M && p.seed_C && !f.mhdr_total;
    auto full_index = [&]() {
        PhFit g = f;
        g.hdr_total = p.seed_hdr_total;
        g.ids_total = p.seed_ids_total;
        if (fits(g)) f = g;
    };
    if (full_first) full_index();
    if (c.hop_slots) {
        PhFit g = f;
        g.hsbits = 32u - c.hop_sshift;
        if (fits(g)) f = g;
    }
    f.clog = 1u;  // two copies of the planes before the full index, the rest after it
    if (!fits(f)) f.clog = 0u;
    if (c.radius < p.M && p.seed_C && !full_first) full_index();
    f.clog = 4u;
    while (f.clog && !fits(f)) f.clog--;
    ph_layout(p.S, bbits, sbits, f);
    return f;
}

// dense ids of a barcode half in the shared-memory key sets: (index1 id + 1) | (index2 id + 1) << 16, 0 in a field =
// not such a key.  key = lo bits | hi bits << length | tag << 40 (PhParams::hop_*)
__device__ __forceinline__ uint32_t ph_hop_find(const PhSmem &sm, const PhParams &c, uint64_t key) {
    const uint32_t w0 = (uint32_t)key, w1 = (uint32_t)(key >> 32);
    const uint32_t h = __umulhi(w0, c.hop_k0) + w0 * c.hop_k1 + w1 * c.hop_k2;
    const uint64_t e = sm.hslots[(h * (uint32_t)sm.hdisp[h >> c.hop_bshift]) >> c.hop_sshift];
    const uint32_t ids = (uint32_t)(e >> 41);  // 11 + 11 bits
    return (e & ((1ull << 41) - 1ull)) == key ? (ids & 0x7FFu) | ((ids >> 11) << 16) : 0u;
}

// metrics.rs:658-676 for an A/C/G/T-only NoMatch read, key sets in shared memory.  Halves of equal length share
// their keys: two look-ups give all four memberships; otherwise four.
__device__ __forceinline__ void ph_hop_check(const DevParams &p, const PhParams &c, const PhSmem &sm, uint32_t lo, uint32_t hi) {
    const uint32_t m1 = (1u << p.i1) - 1u, m2 = (1u << p.i2) - 1u;
    uint32_t a, b, c1, d2;  // ids + 1 of: index1 key in bc[..i1], index2 key in bc[i1..], index1 key in bc[i2..], index2 key in bc[..i2]
    if (p.i1 == p.i2) {
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

// planes of a read held as hash words: compact record (lo | hi << 21) or wide (w0 = lo, w1 = hi)
template <bool WIDE>
__device__ __forceinline__ void ph_planes(uint32_t w0, uint32_t w1, uint32_t &lo, uint32_t &hi) {
    if (WIDE) {
        lo = w0;
        hi = w1;
    } else {
        lo = w0 & kCompactMask;
        hi = __funnelshift_r(w0, w1, kCompactField);  // bits 21..52 of the record; 42..52 are zero for a plain read
    }
}

// the perfect-hash probe of one plain read: entry (sample or S) and distance to it.
// prow = this lane's copy of the planes table, rowbytes = bytes per row of that table.
template <bool WIDE>
__device__ __forceinline__ void ph_probe(const PhSmem &sm, const PhParams &c, const unsigned char *prow, uint32_t rowbytes,
                                         uint32_t w0, uint32_t w1, uint32_t &e, uint32_t &d) {
    const uint32_t h = __umulhi(w0, c.k0) + w0 * c.k1 + w1 * c.k2;
    const uint32_t mul = sm.disp[h >> c.bshift];
    e = sm.slots[(h * mul) >> c.sshift];
    const uint2 t = *reinterpret_cast<const uint2 *>(prow + e * rowbytes);
    uint32_t lo, hi;
    ph_planes<WIDE>(w0, w1, lo, hi);
    d = (uint32_t)__popc((lo ^ t.x) | (hi ^ t.y));
}

// verdict of one read of the general path: patch the placeholder the fast path stored and move its count from
// "Undetermined", where the fast path booked it, to its bin; hop check
__device__ __forceinline__ void ph_emit(const DevParams &p, const PhParams &c, const PhSmem &sm, bool hop_smem, uint32_t S1,
                                        uint64_t r, const Rd &rd, const Verdict &v) {
    if (v.bin != p.S) {
        p.out_idx[r] = (uint16_t)v.bin;
        if (p.out_dist) p.out_dist[r] = (uint8_t)v.dist;
        atomicAdd(&sm.hist[min(v.dist, 2u) * S1 + v.bin], 1u);  // metrics.rs:428-437
        atomicSub(&sm.hist[2u * S1 + p.S], 1u);
    } else if (p.hop_on) {
        if (hop_smem && (rd.nmask | rd.xmask) == 0u) ph_hop_check(p, c, sm, rd.lo, rd.hi);
        else hop_check(p, rd);
    }
}

// VEC: every chunk is full and the three buffers are aligned for the vector accesses (lane i owns consecutive reads).
// !VEC: any alignment, ragged end (lane i owns reads i, i+32, ...).
// SHORT: the key set's radius is smaller than max_mismatches (second level, third queue, seed search of undecided reads).
template <uint32_t MODE, bool VEC, bool WIDE, bool SHORT>
__global__ void __launch_bounds__(kPhThreads, 1) sgdx_match_phx(const DevParams p, const PhParams c, const PhFit fit) {
    constexpr uint32_t RPL = WIDE ? 2u : 4u;  // reads per lane per chunk
    constexpr uint32_t CHUNK = 32u * RPL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhSmem sm;
    const uint32_t bbits = 32u - c.bshift, sbits = 32u - c.sshift;
    sm.planes = reinterpret_cast<uint2 *>(smem_raw);
    sm.hist = reinterpret_cast<uint32_t *>(smem_raw + fit.o_hist);
    sm.disp = reinterpret_cast<uint16_t *>(smem_raw + fit.o_disp);
    sm.slots = reinterpret_cast<uint16_t *>(smem_raw + fit.o_slots);
    sm.qrec = reinterpret_cast<uint64_t *>(smem_raw + fit.o_qrec);
    sm.q2rec = reinterpret_cast<uint64_t *>(smem_raw + fit.o_q2rec);
    sm.q3rec = reinterpret_cast<uint64_t *>(smem_raw + fit.o_q3rec);
    sm.hslots = reinterpret_cast<uint64_t *>(smem_raw + fit.o_hslots);
    sm.qoff = reinterpret_cast<uint32_t *>(smem_raw + fit.o_qoff);
    sm.q2off = reinterpret_cast<uint32_t *>(smem_raw + fit.o_q2off);
    sm.q3off = reinterpret_cast<uint32_t *>(smem_raw + fit.o_q3off);
    sm.hdisp = reinterpret_cast<uint16_t *>(smem_raw + fit.o_hdisp);
    sm.hdr = reinterpret_cast<uint16_t *>(smem_raw + fit.o_hdr);
    sm.ids = reinterpret_cast<uint16_t *>(smem_raw + fit.o_ids);
    sm.mhdr = reinterpret_cast<uint16_t *>(smem_raw + fit.o_mhdr);
    sm.mids = reinterpret_cast<uint16_t *>(smem_raw + fit.o_mids);
    const uint32_t S1 = p.S + 1u, nbins = S1 * 3u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ncopies = 1u << fit.clog;
    const uint32_t lmask = c.lmask;

    copy_table(sm.disp, c.disp, 2u << bbits, threadIdx.x, kPhThreads);
    copy_table(sm.slots, c.slots, 2u << sbits, threadIdx.x, kPhThreads);
    // (the dummy row is >= 11 mismatches away from any compact record)
    fill_planes(sm.planes, p.table, p.S, fit.clog, make_uint2(0u, 0xFFFFFFFFu), threadIdx.x, kPhThreads);
    for (uint32_t i = threadIdx.x; i < nbins; i += kPhThreads) sm.hist[i] = 0u;
    if (fit.hsbits) {
        copy_table(sm.hslots, c.hop_slots, 8u << fit.hsbits, threadIdx.x, kPhThreads);
        copy_table(sm.hdisp, c.hop_disp, 2u << fit.hbbits, threadIdx.x, kPhThreads);
    }
    copy_table(sm.hdr, p.seed_hdr, 2u * fit.hdr_total, threadIdx.x, kPhThreads);
    copy_table(sm.ids, p.seed_ids, 2u * fit.ids_total, threadIdx.x, kPhThreads);
    copy_table(sm.mhdr, c.ms_hdr, 2u * fit.mhdr_total, threadIdx.x, kPhThreads);
    copy_table(sm.mids, c.ms_ids, 2u * fit.mids_total, threadIdx.x, kPhThreads);
    __syncthreads();

    uint64_t *const qrec = sm.qrec + warp * fit.qcap;
    uint32_t *const qoff = sm.qoff + warp * fit.qcap;
    uint64_t *const q2rec = sm.q2rec + warp * kPhQueue2;
    uint32_t *const q2off = sm.q2off + warp * kPhQueue2;
    uint64_t *const q3rec = sm.q3rec + warp * kPhQueue2;
    uint32_t *const q3off = sm.q3off + warp * kPhQueue2;
    const unsigned char *const prow = reinterpret_cast<const unsigned char *>(sm.planes + (lane & (ncopies - 1u)));
    const uint32_t rowbytes = 8u << fit.clog;
    const uint32_t qrec_s = (uint32_t)__cvta_generic_to_shared(qrec), qoff_s = (uint32_t)__cvta_generic_to_shared(qoff);
    const uint16_t *const seed_hdr = fit.hdr_total ? sm.hdr : p.seed_hdr;
    const uint16_t *const seed_ids = fit.ids_total ? sm.ids : p.seed_ids;
    const uint16_t *const ms_hdr = fit.mhdr_total ? sm.mhdr : c.ms_hdr;
    const uint16_t *const ms_ids = fit.mids_total ? sm.mids : c.ms_ids;
    uint32_t n1 = 0u, n2 = 0u, n3 = 0u;  // queue fills, warp-uniform
    const uint64_t total_warps = (uint64_t)gridDim.x * kPhWarps, gwarp = (uint64_t)blockIdx.x * kPhWarps + warp;
    const uint64_t nchunks = (p.n + CHUNK - 1u) / CHUNK;
    const uint32_t below = (1u << lane) - 1u;
    const bool hop = p.hop_on != 0u, hop_smem = fit.hsbits != 0u;
    constexpr bool final_radius = !SHORT;  // radius == max_mismatches: a plain read the probe does not decide is NoMatch
    const uint32_t off0 = VEC ? lane * RPL : lane, offk = VEC ? 1u : 32u;  // chunk offset of my read k = off0 + k * offk

    // The records of a chunk.  Compact: r[k] = record k.  Wide: r[2k] = lo | hi << 32, r[2k+1] = nmask | xmask << 32.
    // Out-of-range reads of the last chunk look invalid; `live` keeps them out of the counters and the queues.
    auto fetch = [&](uint64_t ch, uint64_t (&r)[4]) {
        if (ch >= nchunks) return;
        const uint64_t base = ch * CHUNK;
        if (VEC) {
            const void *src = WIDE ? static_cast<const void *>(c.reads_wide + base + off0) : static_cast<const void *>(c.reads + base + off0);
            asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]) : "l"(src));
        } else {
#pragma unroll
            for (uint32_t k = 0; k < RPL; k++) {
                const uint64_t i = base + k * 32u + lane;
                if (WIDE) {
                    uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
                    if (i < p.n) v = __ldg(c.reads_wide + i);
                    r[2 * k] = (uint64_t)v.x | ((uint64_t)v.y << 32);
                    r[2 * k + 1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
                } else {
                    r[k] = i < p.n ? __ldg(c.reads + i) : ~0ull;
                }
            }
        }
    };
    // wide: does the record carry anything but barcode_len plane bits?
    auto wide_stray = [&](uint64_t planes, uint64_t masks) -> bool {
        return ((uint32_t)masks | (uint32_t)(masks >> 32) | (((uint32_t)planes | (uint32_t)(planes >> 32)) & ~lmask)) != 0u;
    };

    uint64_t rec[4], nextrec[4];
    fetch(gwarp, rec);
    uint32_t iter = 0u;
    for (uint64_t ch = gwarp;; ch += total_warps, iter++) {
        const bool final = ch >= nchunks;
        if (!final) {
            fetch(ch + total_warps, nextrec);
            // 16-byte records: the chunk after the next one is requested into L2 (no registers held; C3 0.910 ->
            // 0.884 ms on one box -- the same request made sgdx_match_ph's compact stream 3 % slower)
            if (VEC && WIDE && ch + 2u * total_warps < nchunks) {
                const uint64_t base2 = (ch + 2u * total_warps) * CHUNK;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(WIDE ? static_cast<const void *>(c.reads_wide + base2 + off0)
                                                                   : static_cast<const void *>(c.reads + base2 + off0)));
            }
            const uint64_t base = ch * CHUNK;
            // the reads of a lane phase by phase, so that their dependent shared-memory chains overlap
            uint32_t bin[RPL], dist[RPL], h[RPL], e[RPL], defer = 0u, strays = 0u;
            uint2 t[RPL];
#pragma unroll
            for (uint32_t k = 0; k < RPL; k++) {
                const uint64_t key = WIDE ? rec[2 * k] : rec[k];
                const uint32_t w0 = (uint32_t)key, w1 = (uint32_t)(key >> 32);
                h[k] = __umulhi(w0, c.k0) + w0 * c.k1 + w1 * c.k2;
                e[k] = sm.disp[h[k] >> c.bshift];
            }
#pragma unroll
            for (uint32_t k = 0; k < RPL; k++) e[k] = sm.slots[(h[k] * e[k]) >> c.sshift];
#pragma unroll
            for (uint32_t k = 0; k < RPL; k++) t[k] = *reinterpret_cast<const uint2 *>(prow + e[k] * rowbytes);
#pragma unroll
            for (uint32_t k = 0; k < RPL; k++) {
                const uint64_t key = WIDE ? rec[2 * k] : rec[k];
                const uint32_t w0 = (uint32_t)key, w1 = (uint32_t)(key >> 32);
                uint32_t lo, hi;
                ph_planes<WIDE>(w0, w1, lo, hi);
                const uint32_t d = (uint32_t)__popc((lo ^ t[k].x) | (hi ^ t[k].y));
                // stray: not "barcode_len A/C/G/T bases" -- an invalid position, or bits above the barcode's length
                const bool stray = WIDE ? wide_stray(rec[2 * k], rec[2 * k + 1]) : ((w0 & c.nm0) | (w1 & c.nm1)) != 0u;
                strays |= stray ? 1u << k : 0u;
                bool bad = d > c.radius || stray;
                if (WIDE) bad = bad || e[k] >= p.S;  // (a compact record cannot come close to the dummy row)
                bin[k] = bad ? p.S : e[k];
                dist[k] = bad ? kNoDist : d;
                // a read that is not decided here is booked as Undetermined for now; its drain moves the count
                const bool live = VEC || base + k * 32u + lane < p.n;
#ifndef SGDX_PH_X_NOHIST  // (experiment builds only, tools/build_variants.sh: what the histogram costs)
                if (live) atomicAdd(&sm.hist[min(dist[k], 2u) * S1 + bin[k]], 1u);
#endif
                // deferred: the hop check of a NoMatch read, or the whole read if it is invalid or undecided
                if (bad && (hop || stray || !final_radius) && live) defer |= 1u << k;
#ifdef SGDX_PH_X_NOQ  // (experiment builds only: what the deferred reads cost)
                defer = 0u;
#endif
            }
            // results: a deferred read gets "Undetermined / no distance" now and is patched later if it matches
            if (VEC) {
                const uint64_t at = base + off0;
                if (WIDE) {
                    *reinterpret_cast<uint32_t *>(p.out_idx + at) = bin[0] | (bin[1] << 16);
                    if (p.out_dist) *reinterpret_cast<uint16_t *>(p.out_dist + at) = (uint16_t)(dist[0] | (dist[1] << 8));
                } else {
                    *reinterpret_cast<uint2 *>(p.out_idx + at) =
                        make_uint2(bin[0] | (bin[1] << 16), bin[2 % RPL] | (bin[3 % RPL] << 16));
                    if (p.out_dist)
                        *reinterpret_cast<uint32_t *>(p.out_dist + at) =
                            dist[0] | (dist[1] << 8) | (dist[2 % RPL] << 16) | (dist[3 % RPL] << 24);
                }
            } else {
#pragma unroll
                for (uint32_t k = 0; k < RPL; k++) {
                    const uint64_t i = base + k * 32u + lane;
                    if (i < p.n) {
                        p.out_idx[i] = (uint16_t)bin[k];
                        if (p.out_dist) p.out_dist[i] = (uint8_t)dist[k];
                    }
                }
            }
            if (__any_sync(0xFFFFFFFFu, defer != 0u)) {
#pragma unroll
                for (uint32_t k = 0; k < RPL; k++) {
                    const bool mine = (defer >> k) & 1u;
                    const uint32_t votes = __ballot_sync(0xFFFFFFFFu, mine);
                    if (mine) {  // wide: the planes travel, flagged if the general path has to fetch the masks too
                        const uint32_t at = n1 + (uint32_t)__popc(votes & below);
                        uint32_t tag = iter * CHUNK + off0 + k * offk;
                        if (WIDE && ((strays >> k) & 1u)) tag |= 0x80000000u;
                        asm volatile("st.shared.u64 [%0], %1;" ::"r"(qrec_s + at * 8u), "l"(WIDE ? rec[2 * k] : rec[k]) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qoff_s + at * 4u), "r"(tag) : "memory");
                    }
                    n1 += (uint32_t)__popc(votes);
                }
                __syncwarp();
            }
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) rec[k] = nextrec[k];
        }

        // ---- deferred reads, 32 at a time.  Plain reads: with the full radius they are NoMatch and only their hop check
        // is left (metrics.rs:658-676); with a short radius they probe the second level here and what that leaves
        // undecided moves on to the second queue (seed search, again 32 at a time).  Reads with invalid positions or
        // stray bits move on to the third queue.
        const auto read_of = [&](uint32_t tag) {  // global read number of a queue entry
            const uint32_t e = tag & 0x7FFFFFFFu;
            return ((uint64_t)(e / CHUNK) * total_warps + gwarp) * CHUNK + (e % CHUNK);
        };
        const auto scan_rest = [&](bool scan, uint64_t r, const Rd &rd) {  // one warp per read over all samples
            uint32_t todo = __ballot_sync(0xFFFFFFFFu, scan);
            while (todo) {
                const uint32_t src = (uint32_t)__ffs(todo) - 1u;
                todo &= todo - 1u;
                const uint64_t qr = __shfl_sync(0xFFFFFFFFu, r, src);
                Rd qrd;
                qrd.lo = __shfl_sync(0xFFFFFFFFu, rd.lo, src);
                qrd.hi = __shfl_sync(0xFFFFFFFFu, rd.hi, src);
                qrd.nmask = __shfl_sync(0xFFFFFFFFu, rd.nmask, src);
                qrd.xmask = __shfl_sync(0xFFFFFFFFu, rd.xmask, src);
                const Verdict qv = warp_scan<MODE>(p, p.table, qrd, lane);
                if (lane == 0u) ph_emit(p, c, sm, hop_smem, S1, qr, qrd, qv);
            }
        };
        const uint2 *const tab = reinterpret_cast<const uint2 *>(prow);
        const auto ms_desc_of = [&](uint32_t cc) { return c.ms_desc[cc]; };
        // the short seed index decides a read (PhParams::ms_*) unless its best candidate is too close to call
        const auto ms_sure = [&](uint32_t best, uint32_t ninv) {
            const uint32_t pb = (best >> 16) - ninv;  // valid-position mismatches of the best candidate
            return best == kInfKey || pb > p.M || 2u * pb + p.delta + ninv < (uint32_t)__ldg(c.nn + (best & 0xFFFFu));
        };
        while (n1 >= 32u || (final && (n1 | n2 | n3))) {
            const uint32_t take = min(n1, 32u);  // 0 when only the other queues are left to flush
            n1 -= take;
            bool to2 = false, to3 = false;
            uint64_t q = 0ull;
            uint32_t off = 0u;
            if (lane < take) {
                q = qrec[n1 + lane];
                off = qoff[n1 + lane];
                const uint32_t w0 = (uint32_t)q, w1 = (uint32_t)(q >> 32);
                to3 = WIDE ? (off >> 31) != 0u : ((w0 & c.nm0) | (w1 & c.nm1)) != 0u;
                if (!to3) {
                    Rd rd = {0u, 0u, 0u, 0u};
                    ph_planes<WIDE>(w0, w1, rd.lo, rd.hi);
                    if (!WIDE) rd.hi &= kCompactMask;
                    if (final_radius) {
                        if (hop_smem) ph_hop_check(p, c, sm, rd.lo, rd.hi);
                        else hop_check(p, rd);
                    } else {
                        to2 = true;
                        if (c.b_slots) {  // second level: the strings within b_radius, in global memory (L2)
                            const uint64_t key = WIDE ? q : ((uint64_t)rd.lo | ((uint64_t)rd.hi << kCompactField));
                            const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
                            const uint32_t hb = __umulhi(k0, c.b_k0) + k0 * c.b_k1 + k1 * c.b_k2;
                            const uint32_t pe = __ldg(c.b_slots + ((hb * (uint32_t)__ldg(c.b_disp + (hb >> c.b_bshift))) >> c.b_sshift));
                            const uint2 tt = *reinterpret_cast<const uint2 *>(prow + pe * rowbytes);
                            const uint32_t pd = (uint32_t)__popc((rd.lo ^ tt.x) | (rd.hi ^ tt.y));
                            Verdict v = {p.S, kNoDist};
                            if (pe < p.S && pd <= c.b_radius) {
                                v.bin = pe;
                                v.dist = pd;
                                to2 = false;
                            } else if (c.b_radius >= p.M) {
                                to2 = false;  // further than max_mismatches from every sample
                            }
                            if (!to2) ph_emit(p, c, sm, hop_smem, S1, read_of(off), rd, v);
                        }
                    }
                }
            }
            if (!final_radius) {
                const uint32_t votes2 = __ballot_sync(0xFFFFFFFFu, to2);
                if (to2) {
                    const uint32_t at = n2 + (uint32_t)__popc(votes2 & below);
                    q2rec[at] = q;
                    q2off[at] = off;
                }
                n2 += (uint32_t)__popc(votes2);
            }
            if (__any_sync(0xFFFFFFFFu, to3)) {
                const uint32_t votes3 = __ballot_sync(0xFFFFFFFFu, to3);
                if (to3) {
                    const uint32_t at = n3 + (uint32_t)__popc(votes3 & below);
                    q3rec[at] = q;
                    q3off[at] = off;
                }
                n3 += (uint32_t)__popc(votes3);
            }
            __syncwarp();
            const bool flush = final && n1 == 0u;

            // ---- second queue (short radius only): plain reads, seed search
            while (!final_radius && (n2 >= 32u || (flush && n2))) {
                const uint32_t takeg = min(n2, 32u);
                n2 -= takeg;
                const bool mine = lane < takeg;
                bool scan = false;
                uint64_t r = 0;
                Rd rd = {0u, 0u, 0u, 0u};
                if (mine) {
                    const uint64_t g = q2rec[n2 + lane];
                    r = read_of(q2off[n2 + lane]);
                    ph_planes<WIDE>((uint32_t)g, (uint32_t)(g >> 32), rd.lo, rd.hi);
                    if (!WIDE) rd.hi &= kCompactMask;
                    uint32_t best = kInfKey, nxt = kInfKey;
                    bool sure = false;
                    if (c.ms_C) {
                        seed_search_t<MODE, false, true>(p, c.ms_C, ms_desc_of, ms_hdr, ms_ids, tab, rd, best, nxt, ncopies);
                        sure = ms_sure(best, 0u);
                        if (!sure) best = nxt = kInfKey;
                    }
                    if (!sure) {
                        if (p.seed_C == 0u) scan = true;
                        else seed_search<MODE, false>(p, seed_hdr, seed_ids, tab, rd, best, nxt, ncopies);
                    }
                    if (!scan) ph_emit(p, c, sm, hop_smem, S1, r, rd, apply_rules<MODE>(p, rd, best, nxt));
                }
                scan_rest(scan, r, rd);
                __syncwarp();
            }

            // ---- third queue: reads with invalid positions or stray bits
            while (n3 >= 32u || (flush && n2 == 0u && n3)) {
                const uint32_t takeg = min(n3, 32u);
                n3 -= takeg;
                const bool mine = lane < takeg;
                bool scan = false;
                uint64_t r = 0;
                Rd rd = {0u, 0u, 0u, 0u};
                if (mine) {
                    const uint32_t tag = q3off[n3 + lane];
                    const uint64_t g = q3rec[n3 + lane];
                    r = read_of(tag);
                    if (WIDE) rd = load_packed(c.reads_wide, r, p.L);  // the masks are not in the queue
                    else rd = rd_of_compact(g, lmask);
                    const uint32_t inv = rd.nmask | rd.xmask;
                    uint32_t best = kInfKey, nxt = kInfKey;
                    if (!cannot_match<MODE>(p, rd)) {
                        if (final_radius && (inv == 0u || (c.single_ok && (inv & (inv - 1u)) == 0u))) {
                            // no invalid base (the record only had stray bits): its own key.  One invalid base: the
                            // unique sample within M over the valid positions, if any, is the entry of one of the 4
                            // completions; whatever else the probes return is a real sample scored at its real distance
                            const uint32_t skip = (MODE == kModePreCompute && rd.nmask != 0u) ? p.M + 1u : 0xFFFFu;
                            const uint32_t nvar = inv ? 4u : 1u;
                            for (uint32_t b = 0; b < nvar; b++) {
                                const uint32_t lo = rd.lo | ((b & 1u) ? inv : 0u), hi = rd.hi | ((b >> 1) ? inv : 0u);
                                const uint64_t key = (uint64_t)lo | ((uint64_t)hi << (WIDE ? 32u : kCompactField));
                                uint32_t pe, pd;
                                ph_probe<WIDE>(sm, c, prow, rowbytes, (uint32_t)key, (uint32_t)(key >> 32), pe, pd);
                                if (pe < p.S) {
                                    const uint2 tt = *reinterpret_cast<const uint2 *>(prow + pe * rowbytes);
                                    score<MODE>(rd, inv, skip, tt.x, tt.y, pe, best, nxt);
                                }
                            }
                        } else {
                            bool sure = false;
                            if (!final_radius && c.ms_C) {
                                sure = seed_search_t<MODE, true, false>(p, c.ms_C, ms_desc_of, ms_hdr, ms_ids, tab, rd, best, nxt, ncopies) &&
                                       ms_sure(best, (uint32_t)__popc(inv));
                                if (!sure) best = nxt = kInfKey;
                            }
                            if (!sure) {
                                // no seed index; or, with the single-N shortcut valid, only reads with several invalid
                                // bases get here: too few to be worth a chain of dependent loads at a lane or two
                                if (p.seed_C == 0u || (final_radius && c.single_ok)) scan = true;
                                else if (!seed_search<MODE, true>(p, seed_hdr, seed_ids, tab, rd, best, nxt, ncopies)) scan = true;
                            }
                        }
                    }
                    if (!scan) ph_emit(p, c, sm, hop_smem, S1, r, rd, apply_rules<MODE>(p, rd, best, nxt));
                }
                scan_rest(scan, r, rd);
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

template <uint32_t MODE, bool VEC, bool WIDE, bool SHORT>
static cudaError_t launch_ph_t(const DevParams &p, const PhParams &c, const PhFit &fit, int sms, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    constexpr uint32_t CHUNK = WIDE ? 64u : 128u;
    const size_t smem = ph_carve(nullptr, p.S, 32u - c.bshift, 32u - c.sshift, fit, nullptr);
    auto kern = sgdx_match_phx<MODE, VEC, WIDE, SHORT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t chunks = (p.n + CHUNK - 1u) / CHUNK;
    uint64_t grid = (chunks + kPhWarps - 1u) / kPhWarps;
    if (grid > (uint64_t)sms) grid = (uint64_t)sms;  // one persistent CTA per SM
    kern<<<(uint32_t)grid, kPhThreads, smem, stream>>>(p, c, fit);
    return cudaGetLastError();
}

template <bool WIDE, bool SHORT>
static cudaError_t launch_ph_w(const DevParams &p0, const PhParams &c0, uint32_t mode, int sms, cudaStream_t stream) {
    DevParams p = p0;
    PhParams c = c0;
    const PhFit fit = ph_fit(p, c);
    const bool ham = mode == kModeHamming;
    const uint64_t chunk = WIDE ? 64u : 128u;
    const void *in = WIDE ? static_cast<const void *>(c.reads_wide) : static_cast<const void *>(c.reads);
    const bool aligned = (reinterpret_cast<uintptr_t>(in) & 31u) == 0u &&
                         (reinterpret_cast<uintptr_t>(p.out_idx) & (WIDE ? 3u : 7u)) == 0u &&
                         (reinterpret_cast<uintptr_t>(p.out_dist) & (WIDE ? 1u : 3u)) == 0u;
    // the full chunks through the vector kernel, the last reads through the general one (a one-CTA launch that loads
    // the tables again: ragged batches of up to 32 M reads take the general kernel alone, see sgdx_ph.cu)
    if (aligned && !(p.n % chunk != 0u && p.n <= (32ull << 20))) {
        const uint64_t n_main = p.n - p.n % chunk;
        p.n = n_main;
        cudaError_t e = ham ? launch_ph_t<kModeHamming, true, WIDE, SHORT>(p, c, fit, sms, stream)
                            : launch_ph_t<kModePreCompute, true, WIDE, SHORT>(p, c, fit, sms, stream);
        if (e != cudaSuccess) return e;
        p.n = p0.n - n_main;
        if (WIDE) c.reads_wide += n_main;
        else c.reads += n_main;
        p.out_idx += n_main;
        if (p.out_dist) p.out_dist += n_main;
    }
    return ham ? launch_ph_t<kModeHamming, false, WIDE, SHORT>(p, c, fit, sms, stream)
               : launch_ph_t<kModePreCompute, false, WIDE, SHORT>(p, c, fit, sms, stream);
}

}  // namespace phx

cudaError_t launch_phx(const DevParams &p, const PhParams &c, uint32_t mode, int sms, cudaStream_t stream) {
    using namespace phx;
    const bool shrt = c.radius < p.M;
    if (c.wide) return shrt ? launch_ph_w<true, true>(p, c, mode, sms, stream) : launch_ph_w<true, false>(p, c, mode, sms, stream);
    return shrt ? launch_ph_w<false, true>(p, c, mode, sms, stream) : launch_ph_w<false, false>(p, c, mode, sms, stream);
}

}  // namespace sgdx
