// sgdx_phbuild.h -- host-side construction of the perfect hash behind sgdx_match_ph (sgdx_ph.cu; PhParams in
// sgdx_kernels.h).  Plain C++ (no CUDA): sgdx_create calls it, and the CPU tests reach it through
// sgdx_ph_probe_host (include/sgdx_synth.h) to check the tables against the oracle without a GPU.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sgdx {

struct PhRules {
    uint32_t S, L, M, delta;
    uint32_t pc_limit;  // PreCompute: furthest visible competitor (DevParams::pc_limit)
    uint32_t mode;      // 0 Hamming rule set, 1 PreCompute rule set
};

struct PhPlanes {
    uint32_t lo, hi;
};

struct PhLevel {  // one perfect hash: keys = the strings within `radius` substitutions of a barcode
    std::vector<uint16_t> disp, slots;
    uint32_t bshift = 0, sshift = 0, k0 = 0, k1 = 0, k2 = 0;
    uint32_t radius = 0, nkeys = 0;
};

struct PhTables {
    PhLevel a;  // first level (shared memory): the largest radius <= max_mismatches that fits 2^15 slots
    PhLevel b;  // second level (global memory; empty when a.radius == max_mismatches or nothing larger fits 2^18 slots)
    uint32_t nm0 = 0, nm1 = 0, lmask = 0, single_ok = 0, dmin = 0;
    uint32_t wide = 0;  // keys are lo | hi << 32 (barcodes of up to 32 bases) instead of compact records
};

struct PhHopTables {
    std::vector<uint16_t> disp;
    std::vector<uint64_t> slots;
    uint32_t bshift = 0, sshift = 0, k0 = 0, k1 = 0, k2 = 0;
};

bool chd_place(const std::vector<uint64_t> &keys, uint32_t bbits, uint32_t sbits, uint32_t max_sbits, uint32_t S,
               size_t smem_limit, size_t (*smem_bytes)(uint32_t, uint32_t, uint32_t), std::vector<uint16_t> *disp,
               std::vector<uint32_t> *where, uint32_t *bshift_out, uint32_t *sshift_out, uint32_t *kset);
// keys = halves as lo bits | hi bits << length (A/C/G/T only), ids = their dense ids; false = not built
bool ph_build_hop(const std::vector<uint64_t> &keys1, const std::vector<uint32_t> &ids1, const std::vector<uint64_t> &keys2,
                  const std::vector<uint32_t> &ids2, bool same_length, PhHopTables *out);

// matcher.rs:175,201: max_mismatches + min_delta - 1 wraps when both are 0 and is then clamped to L
inline uint32_t ph_pc_limit(uint32_t max_mismatches, uint32_t min_delta, uint32_t L) {
    const uint64_t md = (uint64_t)max_mismatches + min_delta;
    uint64_t hi = md == 0 ? L : md - 1;
    if (hi > L) hi = L;
    const uint64_t m = max_mismatches < L ? max_mismatches : L;
    return (uint32_t)(hi > m ? hi : m);
}

// false = this configuration gets no perfect hash (too many keys, tables beyond shared memory, barcode too long)
bool ph_build(const PhRules &r, const std::vector<PhPlanes> &table, bool wide, size_t smem_limit,
              size_t (*smem_bytes)(uint32_t S, uint32_t bbits, uint32_t sbits), PhTables *out);

// the kernel's probe in host arithmetic: entry (sample or S) for a read without invalid positions, given as its key
// (compact record, or lo | hi << 32 for wide tables)
inline uint32_t ph_level_lookup(const PhLevel &t, uint64_t key) {
    const uint32_t w0 = (uint32_t)key, w1 = (uint32_t)(key >> 32);
    const uint32_t h = (uint32_t)(((uint64_t)w0 * t.k0) >> 32) + w0 * t.k1 + w1 * t.k2;
    return t.slots[(h * (uint32_t)t.disp[h >> t.bshift]) >> t.sshift];
}

}  // namespace sgdx
