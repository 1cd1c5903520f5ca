// sgdx_phbuild.cpp -- see sgdx_phbuild.h.  Key set: every A/C/G/T string within max_mismatches substitutions of an
// expected barcode.  Each key's verdict is computed here exactly as the kernels compute it per read (score +
// apply_rules of sgdx_device.cuh for a read without invalid positions): best / second-best (distance << 16 | sample)
// keys over the whole table, then matcher.rs:295-299 or the PreCompute visibility rule.
#include "sgdx_guard.h"
#include "sgdx_phbuild.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "../../include/sgdx.h"
#include "../../include/sgdx_synth.h"

namespace sgdx {

namespace {
constexpr uint32_t kField = 21, kFieldMask = (1u << kField) - 1u, kMaxSlotBits = 15, kInf = 0xFFFFFFFFu;

inline uint64_t record_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << kField); }

uint32_t verdict(const PhRules &r, const std::vector<PhPlanes> &table, uint32_t lo, uint32_t hi) {
    uint32_t best = kInf, nxt = kInf;
    for (uint32_t s = 0; s < r.S; s++) {
        const uint32_t d = (uint32_t)__builtin_popcount((lo ^ table[s].lo) | (hi ^ table[s].hi));
        const uint32_t key = (d << 16) | s;
        nxt = std::min(nxt, std::max(best, key));
        best = std::min(best, key);
    }
    const uint32_t bd = best >> 16, nd = nxt >> 16;  // nd = 0xFFFF without a competitor
    bool ok;
    if (r.mode == 0u) ok = best != kInf && bd <= r.M && (nd - bd) > r.delta;
    else ok = best != kInf && bd <= r.M && !(nd <= bd + r.delta && nd <= r.pc_limit);
    return ok ? (best & 0xFFFFu) : r.S;
}
}  // namespace

namespace {
const uint32_t kChdK[6][3] = {{0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du}, {0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du},
                              {0xFD7046C5u, 0xB55A4F09u, 0x2545F491u}, {0x9E6C63D1u, 0x7FEB352Du, 0x846CA68Bu},
                              {0xCC9E2D51u, 0x1B873593u, 0xE6546B65u}, {0x85EBCA6Bu, 0xC2B2AE35u, 0x9E3779B9u}};
const uint32_t *chd_multipliers(uint32_t set) { return kChdK[set % 6u]; }
}  // namespace

// Two-level perfect hash ("hash, displace") of 64-bit keys:
//     h = hi32(w0 * K0) + w0 * K1 + w1 * K2,   slot = (h * disp[h >> bshift]) >> sshift
// disp holds one odd 16-bit multiplier per bucket; buckets are placed largest first while the table is still empty.
// where[i] = slot of key i.  *kset = index of the multiplier set that worked (chd_multipliers).
bool chd_place(const std::vector<uint64_t> &keys, uint32_t bbits, uint32_t sbits, uint32_t max_sbits, uint32_t S,
               size_t smem_limit, size_t (*smem_bytes)(uint32_t, uint32_t, uint32_t), std::vector<uint16_t> *disp,
               std::vector<uint32_t> *where, uint32_t *bshift_out, uint32_t *sshift_out, uint32_t *kset) {
    const size_t nk = keys.size();
    for (uint32_t attempt = 0; attempt < 12; attempt++) {
        const uint32_t *K = kChdK[attempt % 6u];
        if (attempt == 6u && ++sbits > max_sbits) return false;  // second round: a larger table
        if (smem_bytes && smem_bytes(S, bbits, sbits) > smem_limit) return false;
        const uint32_t nb = 1u << bbits, nslots = 1u << sbits, bshift = 32u - bbits, sshift = 32u - sbits;
        std::vector<uint32_t> hv(nk);
        std::vector<std::vector<uint32_t>> buckets(nb);
        for (size_t i = 0; i < nk; i++) {
            const uint32_t w0 = (uint32_t)keys[i], w1 = (uint32_t)(keys[i] >> 32);
            hv[i] = (uint32_t)(((uint64_t)w0 * K[0]) >> 32) + w0 * K[1] + w1 * K[2];
            buckets[hv[i] >> bshift].push_back((uint32_t)i);
        }
        std::vector<uint32_t> order(nb);
        for (uint32_t b = 0; b < nb; b++) order[b] = b;
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t a, uint32_t b) { return buckets[a].size() > buckets[b].size(); });
        disp->assign(nb, 1u);
        where->assign(nk, 0u);
        std::vector<uint8_t> used(nslots, 0u);
        std::vector<uint32_t> pos;
        bool ok = true;
        for (uint32_t oi = 0; oi < nb && ok; oi++) {
            const std::vector<uint32_t> &bk = buckets[order[oi]];
            if (bk.empty()) break;
            bool placed = false;
            for (uint32_t mul = 1u; mul < 65536u && !placed; mul += 2u) {
                pos.clear();
                bool clash = false;
                for (uint32_t i : bk) {
                    const uint32_t q = (hv[i] * mul) >> sshift;
                    if (used[q] || std::find(pos.begin(), pos.end(), q) != pos.end()) {
                        clash = true;
                        break;
                    }
                    pos.push_back(q);
                }
                if (clash) continue;
                for (size_t k = 0; k < bk.size(); k++) {
                    used[pos[k]] = 1u;
                    (*where)[bk[k]] = pos[k];
                }
                (*disp)[order[oi]] = (uint16_t)mul;
                placed = true;
            }
            ok = placed;
        }
        if (!ok) continue;
        *bshift_out = bshift;
        *sshift_out = sshift;
        *kset = attempt % 6u;
        return true;
    }
    return false;
}

// Hop key sets (metrics.rs:575-614) as one perfect hash.  key = half (lo bits | hi bits << length) | tag << 40, tag = 1
// for an index2 half when the two halves differ in length (with equal lengths a string may be both, and one look-up
// answers both questions); slots[] = key | (index1 id + 1) << 41 | (index2 id + 1) << 52, 0 = empty.
bool ph_build_hop(const std::vector<uint64_t> &keys1, const std::vector<uint32_t> &ids1, const std::vector<uint64_t> &keys2,
                  const std::vector<uint32_t> &ids2, bool same_length, PhHopTables *out) {
    std::vector<std::pair<uint64_t, uint64_t>> kv;  // key -> id fields
    for (size_t i = 0; i < keys1.size(); i++) {
        if (ids1[i] + 1u > 0x7FFu) return false;
        kv.emplace_back(keys1[i], (uint64_t)(ids1[i] + 1u) << 41);
    }
    for (size_t i = 0; i < keys2.size(); i++) {
        if (ids2[i] + 1u > 0x7FFu) return false;
        kv.emplace_back(keys2[i] | (same_length ? 0ull : 1ull << 40), (uint64_t)(ids2[i] + 1u) << 52);
    }
    std::sort(kv.begin(), kv.end());
    std::vector<uint64_t> keys, vals;
    for (const auto &e : kv) {
        if (!keys.empty() && keys.back() == e.first) vals.back() |= e.second;
        else {
            keys.push_back(e.first);
            vals.push_back(e.first | e.second);
        }
    }
    if (keys.empty() || keys.size() > 5000u) return false;  // 64 KB of slots at most
    uint32_t sbits = 4, bbits = 2;
    // load factor <= 0.7; up to 0.8 where that halves a table of 32 KB or more (placement retries with more slots if it fails)
    while ((double)keys.size() > (sbits >= 12u ? 0.8 : 0.7) * (double)(1u << sbits)) sbits++;
    while (((size_t)3 << bbits) < keys.size()) bbits++;
    std::vector<uint32_t> where;
    uint32_t kset = 0;
    if (!chd_place(keys, bbits, sbits, 13u, 0u, 0, nullptr, &out->disp, &where, &out->bshift, &out->sshift, &kset)) return false;
    const uint32_t *K = chd_multipliers(kset);
    out->k0 = K[0];
    out->k1 = K[1];
    out->k2 = K[2];
    out->slots.assign((size_t)1 << (32u - out->sshift), 0ull);
    for (size_t i = 0; i < keys.size(); i++) out->slots[where[i]] = vals[i];
    return true;
}

// keys = the strings within `radius` substitutions of a barcode, entry[i] = verdict of keys[i]
static void ph_enumerate(const PhRules &r, const std::vector<PhPlanes> &table, bool wide, uint32_t radius,
                         std::vector<uint64_t> *keys_out, std::vector<uint16_t> *entry_out) {
    const uint32_t S = r.S, L = r.L;
    auto key_of = [&](uint32_t lo, uint32_t hi) { return wide ? (uint64_t)lo | ((uint64_t)hi << 32) : record_of(lo, hi); };
    struct Node {  // strings at exact distance k from sample s, each once (substituted positions ascending)
        uint32_t lo, hi, next_pos;
    };
    std::vector<uint64_t> &keys = *keys_out;
    keys.clear();
    std::vector<Node> frontier, grown;
    for (uint32_t s = 0; s < S; s++) {
        frontier.assign(1, Node{table[s].lo, table[s].hi, 0u});
        keys.push_back(key_of(table[s].lo, table[s].hi));
        for (uint32_t k = 1; k <= radius; k++) {
            grown.clear();
            for (const Node &nd : frontier)
                for (uint32_t j = nd.next_pos; j < L; j++) {
                    const uint32_t cur = ((nd.lo >> j) & 1u) | (((nd.hi >> j) & 1u) << 1);
                    for (uint32_t b = 0; b < 4; b++) {
                        if (b == cur) continue;
                        const uint32_t lo = (nd.lo & ~(1u << j)) | ((b & 1u) << j);
                        const uint32_t hi = (nd.hi & ~(1u << j)) | ((b >> 1) << j);
                        grown.push_back(Node{lo, hi, j + 1u});
                        keys.push_back(key_of(lo, hi));
                    }
                }
            frontier.swap(grown);
        }
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    entry_out->resize(keys.size());
    for (size_t i = 0; i < keys.size(); i++) {
        const uint32_t lo = wide ? (uint32_t)keys[i] : (uint32_t)keys[i] & kFieldMask;
        const uint32_t hi = wide ? (uint32_t)(keys[i] >> 32) : (uint32_t)(keys[i] >> kField) & kFieldMask;
        (*entry_out)[i] = (uint16_t)verdict(r, table, lo, hi);
    }
}

static double ph_key_estimate(uint32_t S, uint32_t L, uint32_t radius) {  // S * sum_k C(L,k) 3^k, before deduplication
    double per = 0.0, term = 1.0;
    for (uint32_t k = 0; k <= radius; k++) {
        per += term;
        term = term * 3.0 * (double)(L - k) / (double)(k + 1u);
    }
    return per * S;
}

// one level: perfect hash of the radius-neighbourhood into at most 2^max_sbits slots
static bool ph_level(const PhRules &r, const std::vector<PhPlanes> &table, bool wide, uint32_t radius, uint32_t max_sbits,
                     size_t smem_limit, size_t (*smem_bytes)(uint32_t, uint32_t, uint32_t), PhLevel *out) {
    if (ph_key_estimate(r.S, r.L, radius) > 3.0 * (double)(1u << max_sbits)) return false;
    std::vector<uint64_t> keys;
    std::vector<uint16_t> entry;
    ph_enumerate(r, table, wide, radius, &keys, &entry);
    const size_t nk = keys.size();
    uint32_t sbits = 6;
    while ((double)nk > 0.86 * (double)(1u << sbits)) sbits++;
    if (sbits > max_sbits) return false;
    uint32_t bbits = 4;
    while (((size_t)3 << bbits) < nk) bbits++;  // <= 3 keys per bucket on average
    PhLevel t;
    std::vector<uint32_t> where;
    uint32_t kset = 0;
    if (!chd_place(keys, bbits, sbits, max_sbits, r.S, smem_limit, smem_bytes, &t.disp, &where, &t.bshift, &t.sshift, &kset)) return false;
    const uint32_t *K = chd_multipliers(kset);
    t.k0 = K[0];
    t.k1 = K[1];
    t.k2 = K[2];
    t.slots.assign((size_t)1 << (32u - t.sshift), (uint16_t)r.S);
    for (size_t i = 0; i < nk; i++) t.slots[where[i]] = entry[i];
    for (size_t i = 0; i < nk; i++)  // every key reaches its own entry through the kernel's arithmetic
        if (ph_level_lookup(t, keys[i]) != entry[i]) return false;
    t.radius = radius;
    t.nkeys = (uint32_t)nk;
    *out = std::move(t);
    return true;
}

bool ph_build(const PhRules &r, const std::vector<PhPlanes> &table, bool wide, size_t smem_limit,
              size_t (*smem_bytes)(uint32_t, uint32_t, uint32_t), PhTables *out) {
    const uint32_t S = r.S, L = r.L, M = r.M;
    if (L > (wide ? 32u : kField) || S >= 0xFFFFu) return false;
    const uint32_t top = std::min(std::min(M, L), 10u);  // 10: the dummy row of the planes table
    // first level (shared memory): the largest radius whose key set fits the slots
    PhTables t;
    bool have = false;
    for (int radius = (int)top; radius >= 0 && !have; radius--)
        have = ph_level(r, table, wide, (uint32_t)radius, kMaxSlotBits, smem_limit, smem_bytes, &t.a);
    if (!have) return false;
    // second level (global memory, L2 resident): when the first stops short of max_mismatches, the largest radius
    // above it that stays within 2^18 slots
    if (t.a.radius < top)
        for (uint32_t radius = top; radius > t.a.radius; radius--)
            if (ph_level(r, table, wide, radius, 18u, 0, nullptr, &t.b)) break;
    const uint32_t lmask = L >= 32u ? 0xFFFFFFFFu : (1u << L) - 1u;
    t.lmask = lmask;
    t.nm0 = wide ? 0u : ~(lmask | (lmask << kField));
    t.nm1 = wide ? 0u : ~(lmask >> (32u - kField));
    uint32_t dmin = 0xFFFFu;
    for (uint32_t i = 0; i < S && dmin > 0; i++)
        for (uint32_t j = i + 1; j < S; j++)
            dmin = std::min(dmin, (uint32_t)__builtin_popcount((table[i].lo ^ table[j].lo) | (table[i].hi ^ table[j].hi)));
    t.dmin = dmin;
    t.single_ok = (uint64_t)dmin > 2ull * M + r.delta + 1ull ? 1u : 0u;
    t.wide = wide ? 1u : 0u;
    *out = std::move(t);
    return true;
}

}  // namespace sgdx

// Test support (include/sgdx_synth.h): build the tables for a configuration on the host and answer records the way the
// kernel's fast path does.  No device, no matcher: records the kernel defers come back marked.
namespace {
int probe(const sgdx_config *cfg, const uint8_t *barcodes, const uint64_t *compact, const sgdx_read *wide_records, uint64_t n,
          uint16_t *out_idx, uint8_t *out_dist, uint32_t *info) {
    using namespace sgdx;
    const bool wide = wide_records != nullptr;
    if (!cfg || !barcodes || (n && (!(compact || wide_records) || !out_idx || !out_dist))) return SGDX_EINVAL;
    const uint32_t S = cfg->num_samples, L = cfg->barcode_len;
    if (S < 1 || S > SGDX_MAX_SAMPLES || L < 1 || L > (wide ? SGDX_MAX_PACKED_LEN : SGDX_COMPACT_MAX_LEN)) return SGDX_EINVAL;
    std::vector<PhPlanes> table(S);
    for (uint32_t s = 0; s < S; s++) {
        uint32_t lo = 0, hi = 0;
        for (uint32_t j = 0; j < L; j++) {
            const uint32_t c = (barcodes[(size_t)s * L + j] >> 1) & 3u;
            lo |= (c & 1u) << j;
            hi |= (c >> 1) << j;
        }
        table[s] = PhPlanes{lo, hi};
    }
    uint32_t matcher = cfg->matcher;
    if (matcher == SGDX_MATCHER_AUTO) matcher = L <= 12 ? SGDX_MATCHER_PRECOMPUTE : SGDX_MATCHER_CACHED_HAMMING;
    PhRules r{S, L, std::min<uint32_t>(cfg->max_mismatches, 64u), std::min<uint32_t>(cfg->min_delta, 64u),
              ph_pc_limit(cfg->max_mismatches, cfg->min_delta, L), matcher == SGDX_MATCHER_PRECOMPUTE ? 1u : 0u};
    PhTables t;
    const bool have = ph_build(r, table, wide, 0, nullptr, &t);
    if (info) {
        info[0] = have ? 1u : 0u;
        info[1] = t.a.nkeys;
        info[2] = have ? 32u - t.a.bshift : 0u;
        info[3] = have ? 32u - t.a.sshift : 0u;
        info[4] = t.single_ok;
        info[5] = t.dmin;
        info[6] = t.a.radius;
        info[7] = t.b.slots.empty() ? 0u : t.b.radius;
        info[8] = t.b.nkeys;
    }
    if (!have) return SGDX_OK;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t lo, hi;
        bool stray;
        uint64_t key;
        if (wide) {
            const sgdx_read &q = wide_records[i];
            lo = q.lo, hi = q.hi;
            stray = (q.nmask | q.xmask | ((q.lo | q.hi) & ~t.lmask)) != 0u;
            key = (uint64_t)lo | ((uint64_t)hi << 32);
        } else {
            const uint32_t w0 = (uint32_t)compact[i], w1 = (uint32_t)(compact[i] >> 32);
            stray = ((w0 & t.nm0) | (w1 & t.nm1)) != 0u;
            lo = w0 & kFieldMask, hi = (uint32_t)(compact[i] >> kField) & kFieldMask;
            key = compact[i];
        }
        if (stray) {  // general path on the device
            out_idx[i] = 0xFFFFu;
            out_dist[i] = 0xFEu;
            continue;
        }
        // first level, then (for what it leaves undecided) the second, as the kernel does
        bool decided = false;
        uint32_t reach = 0;
        for (const PhLevel *lv : {&t.a, &t.b}) {
            if (lv->slots.empty() || decided) continue;
            const uint32_t e = ph_level_lookup(*lv, key);
            const uint32_t tlo = e < S ? table[e].lo : 0u, thi = e < S ? table[e].hi : 0xFFFFFFFFu;
            const uint32_t d = (uint32_t)__builtin_popcount((lo ^ tlo) | (hi ^ thi));
            reach = lv->radius;
            if (e < S && d <= lv->radius) {
                out_idx[i] = (uint16_t)e;
                out_dist[i] = (uint8_t)d;
                decided = true;
            }
        }
        if (decided) continue;
        if (reach >= r.M) {
            out_idx[i] = (uint16_t)S;
            out_dist[i] = (uint8_t)SGDX_NO_DIST;
        } else {  // undecided: the kernel's general path searches the seed index
            out_idx[i] = 0xFFFFu;
            out_dist[i] = 0xFDu;
        }
    }
    return SGDX_OK;
}
}  // namespace

extern "C" int sgdx_ph_probe_host(const sgdx_config *cfg, const uint8_t *barcodes, const uint64_t *records, uint64_t n,
                                  uint16_t *out_idx, uint8_t *out_dist, uint32_t *info) {
    return sgdx::guarded("sgdx_ph_probe_host", [&]() -> int {
    return probe(cfg, barcodes, records, nullptr, n, out_idx, out_dist, info);
    });
}

extern "C" int sgdx_ph_probe_host_wide(const sgdx_config *cfg, const uint8_t *barcodes, const sgdx_read *records, uint64_t n,
                                       uint16_t *out_idx, uint8_t *out_dist, uint32_t *info) {
    return sgdx::guarded("sgdx_ph_probe_host_wide", [&]() -> int {
    if (n && !records) return SGDX_EINVAL;
    static const sgdx_read none{};
    return probe(cfg, barcodes, nullptr, records ? records : &none, n, out_idx, out_dist, info);
    });
}
