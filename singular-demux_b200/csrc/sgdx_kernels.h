// sgdx_kernels.h -- launcher interface between the C-ABI layer and the kernels.
#pragma once
#include "sgdx_device.cuh"

namespace sgdx {

constexpr int kDirectThreads = 256;
constexpr int kSeededThreads = 256;

size_t direct_smem_bytes(uint32_t S);
cudaError_t launch_direct(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream);
size_t seeded_smem_bytes(const DevParams &p);
size_t exact_smem_bytes(const DevParams &p, uint32_t threads);
uint32_t exact_block_threads(const DevParams &p);
cudaError_t launch_seeded(const DevParams &p, uint32_t mode, bool packed, int sms, cudaStream_t stream);
cudaError_t launch_pack(const uint8_t *reads, uint64_t n, uint32_t L, uint4 *out, int sms, cudaStream_t stream);

// compact records (sgdx_compact.cu, sgdx_ph.cu): one 64-bit word per read of at most kCompactMaxLen bases --
// lo plane | hi plane << 21 | invalid plane << 42 (at an invalid position lo = 0: 'N', lo = 1: any other byte)
constexpr uint32_t kCompactMaxLen = 21;
constexpr uint32_t kCompactField = 21;  // bits per plane
constexpr uint32_t kCompactMask = (1u << kCompactField) - 1u;
inline uint64_t compact_record(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 21); }
cudaError_t launch_pack_compact(const uint8_t *reads, uint64_t n, uint32_t L, uint64_t *out, int sms, cudaStream_t stream);
// dense records (2 L bits per read) + exception list -> compact records (sgdx_unpack_dense, sgdx_patch_exceptions);
// `dense` 4-byte aligned and padded by 32 bytes, `out` 16-byte aligned
cudaError_t launch_unpack_dense(const uint8_t *dense, uint64_t n, uint32_t L, const uint32_t *exc_idx, const uint64_t *exc_rec,
                                uint64_t n_exc, uint64_t *out, int sms, cudaStream_t stream);
cudaError_t launch_unpack_dense_wide(const uint8_t *dense, uint64_t n, uint32_t L, const uint32_t *exc_idx, const uint4 *exc_rec,
                                     uint64_t n_exc, uint4 *out, int sms, cudaStream_t stream);  // 22..32 bases -> sgdx_read records
cudaError_t launch_expand_compact(const uint64_t *in, uint64_t n, uint32_t L, uint4 *out, int sms, cudaStream_t stream);

// Perfect-hash kernel for packed records (sgdx_ph.cu).  Key set K = every A/C/G/T string within `radius`
// substitutions of an expected barcode (radius = max_mismatches when K is small enough, else as large as fits).  For
// a read without invalid positions the verdict is a function of the string alone, so sgdx_create computes it once
// per key (full scan, both rule sets) and stores the winning sample (or "none") in a two-level perfect hash over K:
//     h    = hi32(w0 * k0) + w0 * k1 + w1 * k2          (w0, w1 = the read's planes as two 32-bit words: the compact
//                                                         record lo | hi << 21, or wide: w0 = lo, w1 = hi)
//     slot = (h * disp[h >> bshift]) >> sshift           (disp: one odd 16-bit multiplier per bucket)
//     e    = slots[slot]                                 (sample index; S = none)
// The entry carries no tag: the kernel verifies by distance.  A read inside K finds its own entry, whose sample is
// its best sample at distance <= radius (or S); a read outside K finds some other key's entry, and its distance to
// that sample is > radius because it is further than that from every sample: NoMatch when radius = max_mismatches,
// undecided (general path) otherwise.  Reads with invalid positions (or stray bits) take the general path.
struct PhParams {
    const uint64_t *reads;     // compact records (wide == 0)
    const uint4 *reads_wide;   // sgdx_read records (wide == 1)
    const uint16_t *disp;      // [1 << (32 - bshift)]
    const uint16_t *slots;     // [1 << (32 - sshift)]
    uint32_t bshift, sshift;
    uint32_t k0, k1, k2;
    uint32_t wide;             // keys / records of up to 32 bases
    uint32_t radius;           // <= max_mismatches
    uint32_t nm0, nm1;         // compact: bits of w0 / w1 that a record made of barcode_len A/C/G/T bases never has set
    uint32_t lmask;            // (1 << barcode_len) - 1
    uint32_t single_ok;        // min pairwise table distance > 2M + delta + 1: a read with ONE invalid base is decided by
                               // the unique sample within M over its valid positions, found by probing its 4 completions
    // Second level (optional, global memory / L2; b_slots == nullptr: none): when radius < max_mismatches, the same
    // construction over the strings within b_radius (radius < b_radius <= max_mismatches).  Only the undecided reads
    // of the first level probe it, 32 at a time from their queue; what it leaves undecided searches the seed index.
    const uint16_t *b_disp;
    const uint16_t *b_slots;
    uint32_t b_bshift, b_sshift, b_k0, b_k1, b_k2, b_radius;
    // Short seed index (optional; ms_C == 0: none) for what the levels above leave undecided when radius <
    // max_mismatches: the layout of DevParams' seed index over max_mismatches + 1 chunks instead of T + 1, so with
    // longer keys (<= 6 bases) and shorter buckets.  It enumerates every sample within max_mismatches of the read's
    // valid positions (pigeonhole).  With best = the closest of them at p valid-position mismatches and ninv invalid
    // positions, every other sample is >= nn[best] - ninv - p valid-position mismatches away (triangle inequality over
    // the valid positions), so the enumeration decides the read when no sample is within max_mismatches (NoMatch) or
    // when 2 p + min_delta + ninv < nn[best] (no competitor within min_delta of the best, in either rule set); the few
    // other reads search the full index.
    const uint16_t *ms_hdr;
    const uint16_t *ms_ids;
    uint32_t ms_C, ms_hdr_total, ms_ids_total;
    uint32_t ms_desc[8];
    const uint8_t *nn;         // [S] distance of every barcode to its nearest other barcode (255 for one sample): the
                               // rule above with nn[best] in place of the table-wide minimum dmin
    // hop key sets for shared memory (optional; hop_slots == nullptr: the kernel uses the global tables of DevParams):
    // one perfect hash of the same construction over the distinct index1 / index2 halves (without Undetermined's
    // all-N halves): key = lo bits | hi bits << half length | tag << 40 (tag = 1: an index2 half, only when the two
    // halves differ in length), slot = key | (index1 id + 1) << 41 | (index2 id + 1) << 52, 0 = empty
    const uint16_t *hop_disp;
    const uint64_t *hop_slots;
    uint32_t hop_bshift, hop_sshift, hop_k0, hop_k1, hop_k2;
};
size_t ph_smem_bytes(uint32_t S, uint32_t bbits, uint32_t sbits);
// sgdx_match_ph (sgdx_ph.cu): compact records, radius == max_mismatches.  sgdx_match_phx (sgdx_phx.cu): everything else.
cudaError_t launch_ph(const DevParams &p, const PhParams &c, uint32_t mode, int sms, cudaStream_t stream);
cudaError_t launch_phx(const DevParams &p, const PhParams &c, uint32_t mode, int sms, cudaStream_t stream);

// one expected barcode, no hop checker (sgdx_single.cu): compact records against a single pair of planes
cudaError_t launch_single(const DevParams &p, const uint64_t *reads, uint32_t mode, int sms, cudaStream_t stream);

// barcodes of 33..64 bases (sgdx_long.cu): ASCII input against {loA, hiA, loB, hiB} planes of the expected barcodes
cudaError_t launch_long(const DevParams &p, const uint4 *table, uint32_t mode, int sms, cudaStream_t stream);

// integer-pipe micro-benchmark (sgdx_microbench.cu)
cudaError_t measure_int_peak(int sms, double *popc_per_s, double *lop3_per_s);

}  // namespace sgdx
