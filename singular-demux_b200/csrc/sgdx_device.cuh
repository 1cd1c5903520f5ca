// sgdx_device.cuh -- device-side building blocks shared by the assignment kernels (sm_100a).
//
// Semantics restated from /root/reference/src/lib (citations are file:line in that tree):
//   distance            matcher.rs:335-348      selection + delta rule   matcher.rs:266-303
//   PreCompute rule     matcher.rs:155-261      N gate + sample index    demux.rs:78-80,524-538
//   per-sample counters metrics.rs:428-437      hop check                metrics.rs:658-676
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sgdx {

constexpr uint32_t kModeHamming = 0;     // CachedHammingDistanceMatcher rule set
constexpr uint32_t kModePreCompute = 1;  // PreComputeMatcher rule set
constexpr uint32_t kInfKey = 0xFFFFFFFFu;
constexpr uint32_t kNoDist = 0xFFu;

struct HopSlot {       // open-addressing slot of a barcode-half key set
    uint64_t key;      // (hi plane << 32) | lo plane of the half, bit j = base j of the half
    uint32_t id;       // dense id of the key, 0xFFFFFFFF = empty
    uint32_t pad;
};

struct DevParams {
    const uint8_t *reads_ascii;  // n * L bytes                (ASCII input)
    const uint4 *reads_packed;   // n sgdx_read records         (packed input)
    uint64_t n;
    uint32_t S, L;
    uint32_t M, delta, free_ns;  // clamped to <= 64 by the host (equivalent for L <= 32)
    int32_t max_no_calls;        // -1 = None
    uint32_t pc_limit;           // PreCompute: max(M, M+delta-1) (L if M+delta == 0): furthest visible competitor
    const uint2 *table;          // S entries {lo, hi} planes of the expected barcodes
    uint32_t table_pad;          // S rounded up to a multiple of 2
    uint16_t *out_idx;
    uint8_t *out_dist;           // nullable
    unsigned long long *counters;  // [(S+1)][3] = {other, perfect, one_mismatch}; total = sum of the three
    // hop checker
    uint32_t hop_on, i1, i2;
    uint32_t u1, u2;             // number of distinct index1 / index2 keys (incl. the all-N half)
    uint32_t alln1, alln2;       // dense ids of Undetermined's all-N halves
    const HopSlot *hash1, *hash2;
    uint32_t hmask1, hmask2;     // capacity - 1
    const uint16_t *hop_lut1, *hop_lut2;  // direct-address id tables of the halves (key = lo | hi << len), 0xFFFF = absent
    uint32_t hop_direct;         // 1 when both halves are <= 11 bases and the tables above exist
    unsigned long long *hop_fwd; // [u1][u2]  barcode = I1[a] ++ I2[b]
    unsigned long long *hop_rev; // [u1][u2]  barcode = I2[b] ++ I1[a]   (reversed-orientation rule only)
    // seeded kernel: pigeonhole seed index over the expected barcodes (built by sgdx_create).
    // The L positions are cut into seed_C = T+1 disjoint chunks, T = largest number of valid-position
    // mismatches a sample that matters can have; chunk c is keyed by its first seed_w[c] bases
    // (key = lo bits | hi bits << w).  Bucket `key` of chunk c is the run
    // seed_ids[seed_hdr[off_c + key] .. seed_hdr[off_c + key + 1]) of samples whose barcode has that key.
    const uint16_t *seed_hdr;
    const uint16_t *seed_ids;
    uint32_t seed_C;             // 0 = no seed index (direct kernel only)
    uint32_t seed_hdr_total;     // entries in seed_hdr
    uint32_t seed_ids_total;     // entries in seed_ids (= seed_C * S)
    uint32_t seed_desc[32];      // chunk c: first base of its key window | window width (1..6) << 8 |
                                 //          offset of its bucket headers in seed_hdr << 16
    // exact-match front end of the seeded kernel: an open-addressing table of the expected barcodes keyed
    // by a multiplicative hash of their ASCII words.  A read that equals a barcode byte for byte is a
    // Match with distance 0 without any search iff min pairwise table distance dmin > min_delta
    // (triangle inequality: every other sample is >= dmin away) -- ex_on says so.
    const uint32_t *ex_slots;    // [1 << ex_bits]  (tag << 16) | (sample + 1), 0 = empty; cuckoo: a key sits at
                                 // h >> (32-bits) or ex_second(h) >> (32-bits), so a lookup is two independent probes
    const uint32_t *ex_ascii;    // [S][Lw] barcode bytes as little-endian words, tail bytes 'A'
    uint32_t ex_on, ex_bits;
    uint32_t Lw;                 // ceil(L / 4)
    uint32_t dmin;               // min pairwise Hamming distance of the table (0xFFFF for S == 1)
    uint32_t ex_k[8];            // odd multipliers of the word hash
    // neighbour table (global memory): every string one substitution (A/C/G/T/N) away from a barcode, same
    // word hash.  entry = tag(11) << 21 | letter(3: A C G T N) << 18 | position(5) << 13 | sample(13),
    // 0xFFFFFFFF = empty.  A hit decides the read iff dmin > min_delta + 2 -- nb_on says so.
    const uint32_t *nb_slots;
    uint32_t nb_on, nb_bits;
};

struct Rd {  // one observed barcode as bit planes; bit j = base j, bits >= L are zero
    uint32_t lo, hi, nmask, xmask;
};

__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// finaliser of the exact-match word hash (host builds the table with the same function)
__host__ __device__ inline uint32_t ex_mix(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    return h;
}

// Hash of a barcode held as ASCII words: sum over words of lo32(w*K) + hi32(w*K), then ex_mix.  The high half
// matters: a substitution in the top byte of a word changes lo32(w*K) only in its top 8 bits (256 possible
// differences, so the 4L neighbours of one barcode collide all the time); hi32 keeps the rest.
__host__ __device__ inline uint32_t ex_word_step(uint32_t h, uint32_t w, uint32_t k) {
#ifdef __CUDA_ARCH__
    return __umulhi(w, k) + (w * k + h);  // IMAD + IMAD.HI on the fma pipe
#else
    const uint64_t pr = (uint64_t)w * k;
    return (uint32_t)(pr >> 32) + ((uint32_t)pr + h);
#endif
}

// second cuckoo position of a key with (mixed) hash h; the first one is h >> (32 - bits)
__host__ __device__ inline uint32_t ex_second(uint32_t h) { return h * 0x9E3779B1u + 0x7F4A7C15u; }

// 0x80 in every byte of z that is zero (exact, no cross-byte borrow)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t z) {
    return ~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z | 0x7F7F7F7Fu);
}

// gather bit 0 of each of the 4 bytes of x into a nibble (byte k -> bit k)
__device__ __forceinline__ uint32_t gather4(uint32_t x) {
    return (x * 0x01020408u) >> 24;  // (1<<24)+(1<<17)+(1<<10)+(1<<3): partial products never collide
}

// fold 4 ASCII bases (byte k = base j+k) into the planes.  code = (ascii >> 1) & 3; a byte is a valid base iff it
// equals "ACTG"[code] (PRMT lookup of the four codes of the word); no-calls are found by an exact byte compare.
__device__ __forceinline__ void fold_word(uint32_t w, uint32_t j, Rd &r) {
    const uint32_t c = (w >> 1) & 0x03030303u;
    const uint32_t y = c | (c >> 4);                                     // byte 0: c0 | c1 << 4, byte 2: c2 | c3 << 4
    const uint32_t x = w ^ __byte_perm(0x47544341u, 0u, __byte_perm(y, 0u, 0x4420));  // non-zero byte <=> not A/C/G/T
    const uint32_t bad = ((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) >> 7) & 0x01010101u;
    const uint32_t nflag = zero_bytes(w ^ 0x4E4E4E4Eu) >> 7;             // 'N'
    const uint32_t nn = gather4(nflag);
    r.lo |= (((c & 0x01010101u) * 0x01020408u) >> 24) << j;
    r.hi |= (((c & 0x02020202u) * 0x00810204u) >> 24) << j;
    r.nmask |= nn << j;
    r.xmask |= (gather4(bad) & ~nn) << j;
}

// Load read r (L ASCII bytes) and convert. words_ok: base pointer 4-byte aligned and L % 4 == 0.
__device__ __forceinline__ Rd load_ascii(const uint8_t *__restrict__ reads, uint64_t r, uint32_t L, bool words_ok) {
    Rd rd = {0u, 0u, 0u, 0u};
    const uint8_t *p = reads + r * (uint64_t)L;
    if (words_ok) {
        const uint32_t *pw = reinterpret_cast<const uint32_t *>(p);
        uint32_t nw = L >> 2;
#pragma unroll 8
        for (uint32_t i = 0; i < 8; i++)
            if (i < nw) fold_word(__ldg(pw + i), i * 4, rd);
    } else {
        for (uint32_t j = 0; j < L; j += 4) {
            uint32_t w = 0x41414141u;  // pad with 'A': contributes zero bits to every plane
            for (uint32_t k = 0; k < 4 && j + k < L; k++)
                w = (w & ~(0xFFu << (8 * k))) | ((uint32_t)__ldg(p + j + k) << (8 * k));
            fold_word(w, j, rd);
        }
    }
    // planes of no-call / foreign positions are irrelevant to matching but must not leak into hop keys
    uint32_t inv = rd.nmask | rd.xmask;
    rd.lo &= ~inv;
    rd.hi &= ~inv;
    return rd;
}

__device__ __forceinline__ Rd load_packed(const uint4 *__restrict__ reads, uint64_t r, uint32_t L) {
    uint4 v = __ldg(reads + r);  // one coalesced 128-bit load per read
    uint32_t lm = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
    Rd rd;
    rd.nmask = v.z & lm;
    rd.xmask = v.w & lm & ~rd.nmask;
    uint32_t inv = rd.nmask | rd.xmask;
    rd.lo = v.x & lm & ~inv;
    rd.hi = v.y & lm & ~inv;
    return rd;
}

struct Verdict {
    uint32_t bin;   // sample index, S = Undetermined
    uint32_t dist;  // reported hamming_dist, kNoDist for NoMatch
};

// best / nxt are (distance << 16 | sample) keys over the N-always-mismatch distance p' (Hamming mode) or
// D (PreCompute mode, with invisible competitors already removed); kInfKey = none.
template <uint32_t MODE>
__device__ __forceinline__ Verdict apply_rules(const DevParams &p, const Rd &rd, uint32_t best, uint32_t nxt) {
    Verdict v;
    v.bin = p.S;
    v.dist = kNoDist;
    uint32_t n_no_calls = __popc(rd.nmask);
    bool gated = p.max_no_calls >= 0 && n_no_calls > (uint32_t)p.max_no_calls;  // demux.rs:528-531
    uint32_t bd = best >> 16, nd = nxt >> 16;  // nd = 0xFFFF when there is no competitor
    bool ok;
    uint32_t reported;
    if (MODE == kModeHamming) {
        // matcher.rs:339-341: the first free_ns no-call positions are skipped -> constant discount
        uint32_t d = bd - min(n_no_calls, p.free_ns);
        reported = d;
        ok = best != kInfKey && d <= p.M && (nd - bd) > p.delta;  // matcher.rs:295-299
    } else {
        reported = bd;
        bool rejected = nd <= bd + p.delta && nd <= p.pc_limit;     // visible competitor within delta
        ok = best != kInfKey && rd.xmask == 0u && bd <= p.M && !rejected;
    }
    if (ok && !gated) {
        v.bin = best & 0xFFFFu;
        v.dist = reported;
    }
    return v;
}

__device__ __forceinline__ uint32_t hop_find(const HopSlot *__restrict__ tab, uint32_t mask, uint32_t alln_id,
                                             uint32_t lo, uint32_t hi, uint32_t nm, uint32_t xm, uint32_t len) {
    uint32_t lm = len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u);
    lo &= lm; hi &= lm; nm &= lm; xm &= lm;
    if (xm) return 0xFFFFFFFFu;
    if (nm) return nm == lm ? alln_id : 0xFFFFFFFFu;  // only Undetermined's all-N half contains Ns
    uint64_t key = ((uint64_t)hi << 32) | lo;
    uint32_t i = (uint32_t)mix64(key ^ ((uint64_t)len << 58)) & mask;
    for (;;) {
        HopSlot s = tab[i];
        if (s.id == 0xFFFFFFFFu) return 0xFFFFFFFFu;
        if (s.key == key) return s.id;
        i = (i + 1) & mask;
    }
}

// metrics.rs:658-676 for one NoMatch read (len == i1 + i2 always holds for device batches)
__device__ __forceinline__ void hop_check(const DevParams &p, const Rd &rd) {
    if (p.hop_direct && (rd.nmask | rd.xmask) == 0u) {
        // short halves without invalid bases: four independent table loads instead of up to four hash probes
        const uint32_t m1 = (1u << p.i1) - 1u, m2 = (1u << p.i2) - 1u;
        const uint32_t a = __ldg(p.hop_lut1 + ((rd.lo & m1) | ((rd.hi & m1) << p.i1)));
        const uint32_t b = __ldg(p.hop_lut2 + (((rd.lo >> p.i1) & m2) | (((rd.hi >> p.i1) & m2) << p.i2)));
        const uint32_t c = __ldg(p.hop_lut1 + (((rd.lo >> p.i2) & m1) | (((rd.hi >> p.i2) & m1) << p.i1)));
        const uint32_t d = __ldg(p.hop_lut2 + ((rd.lo & m2) | ((rd.hi & m2) << p.i2)));
        if (a != 0xFFFFu && b != 0xFFFFu) atomicAdd(&p.hop_fwd[(uint64_t)a * p.u2 + b], 1ull);
        else if (c != 0xFFFFu && d != 0xFFFFu) atomicAdd(&p.hop_rev[(uint64_t)c * p.u2 + d], 1ull);
        return;
    }
    uint32_t a = hop_find(p.hash1, p.hmask1, p.alln1, rd.lo, rd.hi, rd.nmask, rd.xmask, p.i1);
    if (a != 0xFFFFFFFFu) {
        uint32_t b = hop_find(p.hash2, p.hmask2, p.alln2, rd.lo >> p.i1, rd.hi >> p.i1, rd.nmask >> p.i1,
                              rd.xmask >> p.i1, p.i2);
        if (b != 0xFFFFFFFFu) {
            atomicAdd(&p.hop_fwd[(uint64_t)a * p.u2 + b], 1ull);
            return;
        }
    }
    // reversed orientation: index1 key in barcode[i2..], index2 key in barcode[..i2]
    uint32_t c = hop_find(p.hash1, p.hmask1, p.alln1, rd.lo >> p.i2, rd.hi >> p.i2, rd.nmask >> p.i2,
                          rd.xmask >> p.i2, p.i1);
    if (c == 0xFFFFFFFFu) return;
    uint32_t d = hop_find(p.hash2, p.hmask2, p.alln2, rd.lo, rd.hi, rd.nmask, rd.xmask, p.i2);
    if (d == 0xFFFFFFFFu) return;
    atomicAdd(&p.hop_rev[(uint64_t)c * p.u2 + d], 1ull);
}

// store the verdict, bump the block-private histogram, run the hop check.  A plain shared-memory atomicAdd of 1
// compiles to ATOMS.POPC.INC, which combines the lanes that hit the same bin in hardware.
__device__ __forceinline__ void emit(const DevParams &p, uint64_t r, bool active, const Rd &rd, const Verdict &v,
                                     uint32_t *hist) {
    if (active) {
        p.out_idx[r] = (uint16_t)v.bin;
        if (p.out_dist) p.out_dist[r] = (uint8_t)v.dist;
        const uint32_t cls = v.dist == 0u ? 1u : (v.dist == 1u ? 2u : 0u);  // metrics.rs:428-437
        atomicAdd(&hist[v.bin * 3u + cls], 1u);
        if (p.hop_on && v.bin == p.S) hop_check(p, rd);
    }
}

}  // namespace sgdx
