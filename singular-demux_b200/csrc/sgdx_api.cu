// sgdx_api.cu -- C-ABI layer (include/sgdx.h): handle, slots/streams, counters, hop export, batcher.
// No torch types, no CPU fallback: every entry point that computes needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "sgdx_guard.h"
#include "sgdx_handle.h"
#include "sgdx_phbuild.h"

using namespace sgdx;

thread_local std::string sgdx::g_err;

int sgdx::fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

extern "C" const char *sgdx_last_error(void) { return g_err.c_str(); }
extern "C" const char *sgdx_version(void) { return "sgdx 0.1.0 (sm_100a)"; }

// Test hooks: environment variables read by sgdx_create that switch an acceleration off (or force one on) so that the
// parity tests can run the slower kernels on the same tables.  They never change results; a build with
// -DSGDX_NO_TEST_HOOKS ignores them.
static inline bool test_hook(const char *name) {
#ifdef SGDX_NO_TEST_HOOKS
    (void)name;
    return false;
#else
    return getenv(name) != nullptr;
#endif
}

static inline uint32_t base_code(uint8_t c) { return (c >> 1) & 3u; }  // A0 C1 T2 G3

static void encode_planes(const uint8_t *s, uint32_t len, uint32_t *lo, uint32_t *hi) {
    uint32_t l = 0, h = 0;
    for (uint32_t j = 0; j < len; j++) {
        uint32_t c = base_code(s[j]);
        l |= (c & 1u) << j;
        h |= (c >> 1) << j;
    }
    *lo = l;
    *hi = h;
}

static void build_hash(const std::vector<std::string> &keys, uint32_t len, std::vector<HopSlot> &tab, uint32_t *mask,
                       uint32_t *alln_id) {
    uint32_t cap = 16;
    while (cap < keys.size() * 2 + 2) cap <<= 1;
    tab.assign(cap, HopSlot{0ull, 0xFFFFFFFFu, 0u});
    *mask = cap - 1;
    *alln_id = 0xFFFFFFFFu;
    for (uint32_t id = 0; id < keys.size(); id++) {
        const std::string &k = keys[id];
        if (k.find('N') != std::string::npos) {  // only Undetermined's all-N half
            *alln_id = id;
            continue;
        }
        uint32_t lo, hi;
        encode_planes(reinterpret_cast<const uint8_t *>(k.data()), len, &lo, &hi);
        uint64_t key = ((uint64_t)hi << 32) | lo;
        uint32_t i = (uint32_t)mix64(key ^ ((uint64_t)len << 58)) & *mask;
        while (tab[i].id != 0xFFFFFFFFu) i = (i + 1) & *mask;
        tab[i] = HopSlot{key, id, 0u};
    }
}

// Pigeonhole seed index (sgdx_seeded.cu): T+1 chunks, bucket headers + sample-id runs.  *have = false
// when the configuration gives no usable index (T+1 > L, or the tables would not fit shared memory).
static int build_seed_index(sgdx_handle *h, const std::vector<uint2> &table, bool *have, double *est_candidates) {
    DevParams &p = h->base;
    *have = false;
    p.seed_C = 0;
    const uint32_t S = p.S, L = p.L;
    // largest number of valid-position mismatches of any sample that can be best or a delta competitor
    const uint64_t T = h->mode == kModeHamming ? (uint64_t)p.M + p.delta : (uint64_t)p.pc_limit;
    if (T + 1 > L) return SGDX_OK;
    const uint32_t C = (uint32_t)T + 1u;
    const uint32_t wmax = C <= 8 ? 5u : 4u;  // keeps the headers <= 32 KB
    if ((uint64_t)C * S > 0xFFFFu) return SGDX_OK;  // bucket bounds are 16-bit offsets into the id runs
    uint32_t total = 0;
    double est = 0.0;
    std::vector<uint32_t> pos(C), w(C), off(C);
    for (uint32_t c = 0; c < C; c++) {
        uint32_t a = (uint32_t)((uint64_t)c * L / C), b = (uint32_t)((uint64_t)(c + 1) * L / C);
        pos[c] = a;
        w[c] = std::min(b - a, wmax);
        off[c] = total;
        total += (1u << (2 * w[c])) + 1u;  // bucket k of chunk c is ids[hdr[off+k] .. hdr[off+k+1])
        est += (double)S / (double)(1u << (2 * w[c]));
        p.seed_desc[c] = pos[c] | (w[c] << 8) | (off[c] << 16);
    }
    std::vector<uint16_t> hdr(total, 0);
    std::vector<uint16_t> ids((size_t)C * S, 0);
    for (uint32_t c = 0; c < C; c++) {
        const uint32_t wm = (1u << w[c]) - 1u, nb = 1u << (2 * w[c]);
        auto key_of = [&](uint32_t s) { return ((table[s].x >> pos[c]) & wm) | (((table[s].y >> pos[c]) & wm) << w[c]); };
        std::vector<uint32_t> cnt(nb, 0u);
        for (uint32_t s = 0; s < S; s++) cnt[key_of(s)]++;
        uint32_t run = c * S;
        for (uint32_t k = 0; k < nb; k++) {
            hdr[off[c] + k] = (uint16_t)run;
            run += cnt[k];
            cnt[k] = 0;
        }
        hdr[off[c] + nb] = (uint16_t)run;
        for (uint32_t s = 0; s < S; s++) {  // ascending sample order inside a bucket
            uint32_t k = key_of(s);
            ids[hdr[off[c] + k] + cnt[k]++] = (uint16_t)s;
        }
    }
    p.seed_hdr_total = total;
    p.seed_ids_total = C * S;
    if (seeded_smem_bytes(p) > 100u * 1024u) {  // keep >= 2 blocks per SM; larger tables stay on the direct kernel
        p.seed_hdr_total = p.seed_ids_total = 0;
        return SGDX_OK;
    }
    CU(cudaMalloc(&h->d_seed_hdr, hdr.size() * sizeof(uint16_t)));
    CU(cudaMalloc(&h->d_seed_ids, ids.size() * sizeof(uint16_t)));
    CU(cudaMemcpy(h->d_seed_hdr, hdr.data(), hdr.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_seed_ids, ids.data(), ids.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    p.seed_hdr = h->d_seed_hdr;
    p.seed_ids = h->d_seed_ids;
    p.seed_C = C;
    *have = true;
    *est_candidates = est;
    return SGDX_OK;
}

// Two-choice cuckoo insertion (random walk).  owner[] remembers the hash of the entry in each slot so that an
// evicted entry can find its other position.  false = gave up (caller grows the table).
static bool cuckoo_insert(std::vector<uint32_t> &slots, std::vector<uint32_t> &owner, uint32_t empty, uint32_t bits,
                          uint32_t hsh, uint32_t entry) {
    uint32_t last = 0xFFFFFFFFu;
    for (int kick = 0; kick < 2000; kick++) {
        const uint32_t p1 = hsh >> (32u - bits), p2 = ex_second(hsh) >> (32u - bits);
        for (uint32_t pos : {p1, p2})
            if (slots[pos] == empty) {
                slots[pos] = entry;
                owner[pos] = hsh;
                return true;
            }
        const uint32_t victim = p1 == last ? p2 : p1;
        std::swap(entry, slots[victim]);
        std::swap(hsh, owner[victim]);
        last = victim;
    }
    return false;
}

// Exact-match front end of the seeded kernel (sgdx_seeded.cu): barcode rows as padded ASCII words and
// an open-addressing table over them.  Enabled only when a byte-exact hit decides the read on its own:
// min pairwise distance of the table > min_delta.
static int build_exact_table(sgdx_handle *h, const uint8_t *barcodes, const std::vector<uint2> &table,
                             double est_candidates) {
    DevParams &p = h->base;
    const uint32_t S = p.S, L = p.L;
    p.ex_on = 0;
    p.Lw = (L + 3u) / 4u;
    uint32_t dmin = 0xFFFFu;
    for (uint32_t i = 0; i < S && dmin > 0; i++)
        for (uint32_t j = i + 1; j < S; j++) {
            uint32_t d = (uint32_t)__builtin_popcount((table[i].x ^ table[j].x) | (table[i].y ^ table[j].y));
            if (d < dmin) dmin = d;
        }
    p.dmin = dmin;
    static const uint32_t K[8] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu,
                                  0x165667B1u, 0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u};
    for (int i = 0; i < 8; i++) p.ex_k[i] = K[i];
    uint32_t bits = 6;
    while ((1u << bits) < 4u * S) bits++;
    p.ex_bits = bits;
    if (test_hook("SGDX_DISABLE_EXACT")) return SGDX_OK;          // test hook: plain seeded kernel
    if (dmin <= p.delta) return SGDX_OK;                       // an exact hit could still be delta-rejected
    if (exact_block_threads(p) == 0u) return SGDX_OK;          // tables that do not fit one SM: general path only
    std::vector<uint32_t> rows((size_t)S * p.Lw, 0x41414141u), slots((size_t)1 << bits, 0u), owner((size_t)1 << bits, 0u);
    for (uint32_t s = 0; s < S; s++) {
        memcpy(&rows[(size_t)s * p.Lw], barcodes + (size_t)s * L, L);  // little-endian host; tail bytes stay 'A'
        uint32_t hsh = 0;
        for (uint32_t i = 0; i < p.Lw; i++) hsh = ex_word_step(hsh, rows[(size_t)s * p.Lw + i], K[i]);
        hsh = ex_mix(hsh);
        // load factor <= 1/4: two-choice cuckoo placement does not fail in practice; if it does, no front end
        if (!cuckoo_insert(slots, owner, 0u, bits, hsh, ((hsh & 0xFFFFu) << 16) | (s + 1u))) return SGDX_OK;
    }
    CU(cudaMalloc(&h->d_ex_slots, slots.size() * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_ex_ascii, rows.size() * sizeof(uint32_t)));
    CU(cudaMemcpy(h->d_ex_slots, slots.data(), slots.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_ex_ascii, rows.data(), rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    p.ex_slots = h->d_ex_slots;
    p.ex_ascii = h->d_ex_ascii;
    p.ex_on = 1;

    // neighbour table: a string at distance 1 from a barcode decides the read when every other barcode is
    // then still further than min_delta away from it: (dmin - 1) - 1 > delta
    p.nb_on = 0;
    if (test_hook("SGDX_DISABLE_NEIGHBOURS")) return SGDX_OK;  // test hook: exercise the seed search on every miss
    if ((uint64_t)dmin <= (uint64_t)p.delta + 2u || S > 0x1FFFu) return SGDX_OK;
    // The neighbour stage costs two L2 probes per missed read; it pays when the seed search it saves is long
    // (measured on a B200: C3, ~12 expected candidates per read: 2.11 ms with, 2.82 ms without per 100 M reads;
    // C2, ~1.5 candidates: 1.65 ms with, 1.56 ms without)
    if (est_candidates < 4.0 && !test_hook("SGDX_FORCE_NEIGHBOURS")) return SGDX_OK;
    const uint64_t entries = (uint64_t)S * L * 4u;
    uint32_t nbits = 8;
    while ((1ull << nbits) < 4u * entries) nbits++;
    if (nbits > 26) return SGDX_OK;
    static const uint8_t letters[5] = {'A', 'C', 'G', 'T', 'N'};
    std::vector<uint32_t> nb, nb_owner, row(p.Lw);
    bool placed = false;
    for (int attempt = 0; attempt < 2 && !placed && nbits <= 26; attempt++, nbits += placed ? 0 : 1) {
        nb.assign((size_t)1 << nbits, 0xFFFFFFFFu);
        nb_owner.assign((size_t)1 << nbits, 0u);
        placed = true;
        for (uint32_t s = 0; s < S && placed; s++)
            for (uint32_t pos = 0; pos < L && placed; pos++)
                for (uint32_t l = 0; l < 5 && placed; l++) {
                    if (letters[l] == barcodes[(size_t)s * L + pos]) continue;
                    for (uint32_t i = 0; i < p.Lw; i++) row[i] = rows[(size_t)s * p.Lw + i];
                    reinterpret_cast<uint8_t *>(row.data())[pos] = letters[l];
                    uint32_t hsh = 0;
                    for (uint32_t i = 0; i < p.Lw; i++) hsh = ex_word_step(hsh, row[i], K[i]);
                    hsh = ex_mix(hsh);
                    placed = cuckoo_insert(nb, nb_owner, 0xFFFFFFFFu, nbits, hsh,
                                           ((hsh & 0x7FFu) << 21) | (l << 18) | (pos << 13) | s);
                }
    }
    if (!placed) return SGDX_OK;  // not seen with load factor <= 1/4; stage 1 then just forwards to the seed search
    CU(cudaMalloc(&h->d_nb_slots, nb.size() * sizeof(uint32_t)));
    CU(cudaMemcpy(h->d_nb_slots, nb.data(), nb.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    p.nb_slots = h->d_nb_slots;
    p.nb_bits = nbits;
    p.nb_on = 1;
    return SGDX_OK;
}

// Perfect hash of the match neighbourhood for sgdx_match_ph (sgdx_ph.cu): built on the host by sgdx_phbuild.cpp
static int build_ph_table(sgdx_handle *h, const std::vector<uint2> &table) {
    const DevParams &p = h->base;
    h->ph = PhParams{};
    if (test_hook("SGDX_DISABLE_PH")) return SGDX_OK;  // test hook: the older compact kernels
    std::vector<PhPlanes> planes(p.S);
    for (uint32_t s = 0; s < p.S; s++) planes[s] = PhPlanes{table[s].x, table[s].y};
    const PhRules rules{p.S, p.L, p.M, p.delta, p.pc_limit, h->mode};
    PhTables t;
    // barcodes of at most 21 bases: keys are compact records; longer ones: keys of the 16-byte sgdx_read records
    if (!ph_build(rules, planes, p.L > kCompactMaxLen, 227u * 1024u, ph_smem_bytes, &t)) return SGDX_OK;
    auto upload = [](const std::vector<uint16_t> &v, uint16_t **d) -> int {
        CU(cudaMalloc(d, v.size() * sizeof(uint16_t)));
        CU(cudaMemcpy(*d, v.data(), v.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        return SGDX_OK;
    };
    if (int rc = upload(t.a.disp, &h->d_ph_disp)) return rc;
    if (int rc = upload(t.a.slots, &h->d_ph_slots)) return rc;
    h->ph.disp = h->d_ph_disp;
    h->ph.slots = h->d_ph_slots;
    h->ph.bshift = t.a.bshift;
    h->ph.sshift = t.a.sshift;
    h->ph.k0 = t.a.k0;
    h->ph.k1 = t.a.k1;
    h->ph.k2 = t.a.k2;
    h->ph.radius = t.a.radius;
    if (!t.b.slots.empty()) {
        if (int rc = upload(t.b.disp, &h->d_ph_b_disp)) return rc;
        if (int rc = upload(t.b.slots, &h->d_ph_b_slots)) return rc;
        h->ph.b_disp = h->d_ph_b_disp;
        h->ph.b_slots = h->d_ph_b_slots;
        h->ph.b_bshift = t.b.bshift;
        h->ph.b_sshift = t.b.sshift;
        h->ph.b_k0 = t.b.k0;
        h->ph.b_k1 = t.b.k1;
        h->ph.b_k2 = t.b.k2;
        h->ph.b_radius = t.b.radius;
    }
    if (t.a.radius < p.M && (uint64_t)p.M + 1u <= p.L && p.M + 1u <= 8u && (uint64_t)(p.M + 1u) * p.S <= 0xFFFFu) {
        // short seed index (PhParams::ms_*): max_mismatches + 1 chunks keyed by their first <= 6 bases
        const uint32_t C = p.M + 1u, S = p.S, L = p.L;
        const uint32_t wmax = C <= 4u ? 6u : 5u;  // headers <= 32 KB
        std::vector<uint32_t> pos(C), w(C), off(C);
        uint32_t total = 0;
        for (uint32_t c = 0; c < C; c++) {
            const uint32_t a = (uint32_t)((uint64_t)c * L / C), b = (uint32_t)((uint64_t)(c + 1) * L / C);
            pos[c] = a;
            w[c] = std::min(b - a, wmax);
            off[c] = total;
            total += (1u << (2 * w[c])) + 1u;
            h->ph.ms_desc[c] = pos[c] | (w[c] << 8) | (off[c] << 16);
        }
        std::vector<uint16_t> hdr(total, 0), ids((size_t)C * S, 0);
        for (uint32_t c = 0; c < C; c++) {
            const uint32_t wm = (1u << w[c]) - 1u, nb = 1u << (2 * w[c]);
            auto key_of = [&](uint32_t s) { return ((table[s].x >> pos[c]) & wm) | (((table[s].y >> pos[c]) & wm) << w[c]); };
            std::vector<uint32_t> cnt(nb, 0u);
            for (uint32_t s = 0; s < S; s++) cnt[key_of(s)]++;
            uint32_t run = c * S;
            for (uint32_t k = 0; k < nb; k++) {
                hdr[off[c] + k] = (uint16_t)run;
                run += cnt[k];
                cnt[k] = 0;
            }
            hdr[off[c] + nb] = (uint16_t)run;
            for (uint32_t s = 0; s < S; s++) {
                const uint32_t k = key_of(s);
                ids[hdr[off[c] + k] + cnt[k]++] = (uint16_t)s;
            }
        }
        if (total < 0x10000u) {  // chunk descriptors carry 16-bit header offsets
            std::vector<uint8_t> nn(S, 255);  // distance to the nearest other barcode
            for (uint32_t i = 0; i < S; i++)
                for (uint32_t j = i + 1; j < S; j++) {
                    const uint8_t d = (uint8_t)__builtin_popcount((table[i].x ^ table[j].x) | (table[i].y ^ table[j].y));
                    nn[i] = std::min(nn[i], d);
                    nn[j] = std::min(nn[j], d);
                }
            CU(cudaMalloc(&h->d_ph_nn, S));
            CU(cudaMemcpy(h->d_ph_nn, nn.data(), S, cudaMemcpyHostToDevice));
            h->ph.nn = h->d_ph_nn;
            if (int rc = upload(hdr, &h->d_ph_ms_hdr)) return rc;
            if (int rc = upload(ids, &h->d_ph_ms_ids)) return rc;
            h->ph.ms_hdr = h->d_ph_ms_hdr;
            h->ph.ms_ids = h->d_ph_ms_ids;
            h->ph.ms_C = C;
            h->ph.ms_hdr_total = total;
            h->ph.ms_ids_total = C * S;
        }
    }
    h->ph.nm0 = t.nm0;
    h->ph.nm1 = t.nm1;
    h->ph.lmask = t.lmask;
    h->ph.single_ok = t.single_ok;
    h->ph.wide = t.wide;
    h->ph_keys = t.a.nkeys;
    if (p.hop_on) {
        // hop key sets for the kernel's shared memory (PhParams::hop_*): the distinct halves without the all-N ones of
        // Undetermined (a read with a no-call takes the general path and the global tables)
        std::vector<uint64_t> k1, k2;
        std::vector<uint32_t> id1, id2;
        auto collect = [](const std::vector<std::string> &keys, uint32_t len, std::vector<uint64_t> &k, std::vector<uint32_t> &id) {
            for (uint32_t i = 0; i < keys.size(); i++) {
                if (keys[i].find('N') != std::string::npos) continue;
                uint32_t lo, hi;
                encode_planes(reinterpret_cast<const uint8_t *>(keys[i].data()), len, &lo, &hi);
                k.push_back((uint64_t)lo | ((uint64_t)hi << len));
                id.push_back(i);
            }
        };
        collect(h->keys1, p.i1, k1, id1);
        collect(h->keys2, p.i2, k2, id2);
        PhHopTables ht;
        if (ph_build_hop(k1, id1, k2, id2, p.i1 == p.i2, &ht)) {
            CU(cudaMalloc(&h->d_ph_hop_slots, ht.slots.size() * sizeof(uint64_t)));
            CU(cudaMalloc(&h->d_ph_hop_disp, ht.disp.size() * sizeof(uint16_t)));
            CU(cudaMemcpy(h->d_ph_hop_slots, ht.slots.data(), ht.slots.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(h->d_ph_hop_disp, ht.disp.data(), ht.disp.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
            h->ph.hop_slots = h->d_ph_hop_slots;
            h->ph.hop_disp = h->d_ph_hop_disp;
            h->ph.hop_bshift = ht.bshift;
            h->ph.hop_sshift = ht.sshift;
            h->ph.hop_k0 = ht.k0;
            h->ph.hop_k1 = ht.k1;
            h->ph.hop_k2 = ht.k2;
        }
    }
    return SGDX_OK;
}

// everything sgdx_create does once the arguments are validated; on failure the caller destroys the handle
static int fill_handle(sgdx_handle *h, const sgdx_config *cfg, const uint8_t *barcodes) {
    const uint32_t S = cfg->num_samples, L = cfg->barcode_len;
    h->cfg = *cfg;
    if (h->cfg.matcher == SGDX_MATCHER_AUTO)  // run.rs:196-198
        h->cfg.matcher = L <= 12 ? SGDX_MATCHER_PRECOMPUTE : SGDX_MATCHER_CACHED_HAMMING;
    if (h->cfg.slots == 0) h->cfg.slots = 2;
    if (h->cfg.slots > 64) h->cfg.slots = 64;
    h->mode = h->cfg.matcher == SGDX_MATCHER_PRECOMPUTE ? kModePreCompute : kModeHamming;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    h->sms = prop.multiProcessorCount;

    DevParams &p = h->base;
    p.S = S;
    p.L = L;
    // distances never exceed L <= 32, so clamping the thresholds at 64 changes no verdict and keeps
    // the 16-bit distance field of the keys safe
    p.M = std::min<uint32_t>(cfg->max_mismatches, 64u);
    p.delta = std::min<uint32_t>(cfg->min_delta, 64u);
    p.free_ns = std::min<uint32_t>(cfg->free_ns, 64u);
    p.max_no_calls = cfg->max_no_calls < 0 ? -1 : std::min<int32_t>(cfg->max_no_calls, 64);
    {
        // matcher.rs:175: `max_mismatches + min_delta - 1` wraps when both are 0 and is then clamped to L (201)
        uint64_t md = (uint64_t)cfg->max_mismatches + cfg->min_delta;
        uint64_t hi = md == 0 ? L : md - 1;
        if (hi > L) hi = L;
        p.pc_limit = (uint32_t)std::max<uint64_t>(hi, std::min<uint32_t>(cfg->max_mismatches, L));
    }
    p.table_pad = (S + 1u) & ~1u;
    p.hop_on = cfg->index1_len > 0 ? 1u : 0u;
    p.i1 = cfg->index1_len;
    p.i2 = cfg->index2_len;

    const bool is_long = L > SGDX_MAX_PACKED_LEN;  // two halves of planes, sgdx_match_long only (sgdx_long.cu)
    std::vector<uint2> table(p.table_pad, make_uint2(0u, 0u));
    if (is_long) {
        std::vector<uint4> tl(S);
        for (uint32_t s = 0; s < S; s++) {
            const uint8_t *bc = barcodes + (uint64_t)s * L;
            encode_planes(bc, 32u, &tl[s].x, &tl[s].y);
            encode_planes(bc + 32, L - 32u, &tl[s].z, &tl[s].w);
        }
        CU(cudaMalloc(&h->d_table_long, tl.size() * sizeof(uint4)));
        CU(cudaMemcpy(h->d_table_long, tl.data(), tl.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    } else {
        for (uint32_t s = 0; s < S; s++) encode_planes(barcodes + (uint64_t)s * L, L, &table[s].x, &table[s].y);
        CU(cudaMalloc(&h->d_table, table.size() * sizeof(uint2)));
        CU(cudaMemcpy(h->d_table, table.data(), table.size() * sizeof(uint2), cudaMemcpyHostToDevice));
        p.table = h->d_table;
    }

    const size_t ncounters = (size_t)(S + 1) * 3;
    CU(cudaMalloc(&h->d_counters, ncounters * sizeof(unsigned long long)));
    CU(cudaMemset(h->d_counters, 0, ncounters * sizeof(unsigned long long)));
    p.counters = h->d_counters;

    if (p.hop_on) {
        // SampleBarcodeHopChecker::try_new (metrics.rs:596-609): halves of every sample AND of Undetermined
        std::map<std::string, uint32_t> m1, m2;
        auto add = [](std::map<std::string, uint32_t> &m, std::vector<std::string> &v, const std::string &k) {
            if (m.emplace(k, (uint32_t)v.size()).second) v.push_back(k);
        };
        for (uint32_t s = 0; s <= S; s++) {
            std::string bc = s < S ? std::string(reinterpret_cast<const char *>(barcodes) + (uint64_t)s * L, L)
                                   : std::string(L, 'N');
            add(m1, h->keys1, bc.substr(0, p.i1));
            add(m2, h->keys2, bc.substr(p.i1));
        }
        p.u1 = (uint32_t)h->keys1.size();
        p.u2 = (uint32_t)h->keys2.size();
        std::vector<HopSlot> t1, t2;
        build_hash(h->keys1, p.i1, t1, &p.hmask1, &p.alln1);
        build_hash(h->keys2, p.i2, t2, &p.hmask2, &p.alln2);
        CU(cudaMalloc(&h->d_hash1, t1.size() * sizeof(HopSlot)));
        CU(cudaMalloc(&h->d_hash2, t2.size() * sizeof(HopSlot)));
        CU(cudaMemcpy(h->d_hash1, t1.data(), t1.size() * sizeof(HopSlot), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_hash2, t2.data(), t2.size() * sizeof(HopSlot), cudaMemcpyHostToDevice));
        p.hash1 = h->d_hash1;
        p.hash2 = h->d_hash2;
        if (p.i1 <= 11u && p.i2 <= 11u && p.u1 < 0xFFFFu && p.u2 < 0xFFFFu) {  // direct-address tables, <= 8 MB each
            auto build_lut = [&](const std::vector<std::string> &keys, uint32_t len, uint16_t **d_out) -> int {
                std::vector<uint16_t> lut((size_t)1 << (2 * len), 0xFFFFu);
                for (uint32_t id = 0; id < keys.size(); id++) {
                    if (keys[id].find('N') != std::string::npos) continue;  // Undetermined's half: reads with N take the hash path
                    uint32_t lo, hi;
                    encode_planes(reinterpret_cast<const uint8_t *>(keys[id].data()), len, &lo, &hi);
                    lut[lo | (hi << len)] = (uint16_t)id;
                }
                CU(cudaMalloc(d_out, lut.size() * sizeof(uint16_t)));
                CU(cudaMemcpy(*d_out, lut.data(), lut.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
                return SGDX_OK;
            };
            if (int rc = build_lut(h->keys1, p.i1, &h->d_hop_lut1)) return rc;
            if (int rc = build_lut(h->keys2, p.i2, &h->d_hop_lut2)) return rc;
            p.hop_lut1 = h->d_hop_lut1;
            p.hop_lut2 = h->d_hop_lut2;
            p.hop_direct = 1;
        }
        size_t cells = (size_t)p.u1 * p.u2;
        CU(cudaMalloc(&h->d_hop_fwd, cells * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_hop_rev, cells * sizeof(unsigned long long)));
        CU(cudaMemset(h->d_hop_fwd, 0, cells * sizeof(unsigned long long)));
        CU(cudaMemset(h->d_hop_rev, 0, cells * sizeof(unsigned long long)));
        p.hop_fwd = h->d_hop_fwd;
        p.hop_rev = h->d_hop_rev;
    }

    // kernel choice: the seeded kernel needs its index to exist (T+1 <= L, shared memory fits); AUTO takes
    // it when the expected number of candidates per read is well below a full scan of the table
    if (is_long) {
        h->seeded = false;
        h->cfg.kernel = SGDX_KERNEL_DIRECT;  // (the long form of the per-pair scan)
    } else {
        double est = 0.0;
        bool have = false;
        if (h->cfg.kernel != SGDX_KERNEL_DIRECT) {
            if (int rc = build_seed_index(h, table, &have, &est)) {
                return rc;
            }
        }
        if (h->cfg.kernel == SGDX_KERNEL_AUTO) h->seeded = have && 2.0 * est + 16.0 < (double)S;
        else h->seeded = have && h->cfg.kernel == SGDX_KERNEL_SEEDED;
        if (h->seeded)
            if (int rc = build_exact_table(h, barcodes, table, est)) {
                return rc;
            }
        h->cfg.kernel = h->seeded ? SGDX_KERNEL_SEEDED : SGDX_KERNEL_DIRECT;
        if (int rc = build_ph_table(h, table)) return rc;
        // one expected barcode and no hop checker: compact records need no table at all (sgdx_single.cu);
        // SGDX_DISABLE_SINGLE is a test hook (the perfect-hash kernel serves the same records)
        h->single = S == 1 && !p.hop_on && L <= kCompactMaxLen && !test_hook("SGDX_DISABLE_SINGLE");
    }

    h->slots.resize(h->cfg.slots);
    for (Slot &s : h->slots) {
        CU(cudaStreamCreateWithFlags(&s.own, cudaStreamNonBlocking));
        s.stream = s.own;
        CU(cudaEventCreate(&s.e0));
        CU(cudaEventCreate(&s.e1));
    }
    // table uploads and memsets above ran on the legacy default stream; the slots' streams are non-blocking and
    // do not order against it, so make everything resident before the first batch can be queued
    CU(cudaDeviceSynchronize());
    return SGDX_OK;
}


extern "C" int sgdx_create(const sgdx_config *cfg, const uint8_t *barcodes, sgdx_handle **out) {
    return sgdx::guarded("sgdx_create", [&]() -> int {
    if (!cfg || !barcodes || !out) return fail(SGDX_EINVAL, "sgdx_create: null argument");
    *out = nullptr;
    const uint32_t S = cfg->num_samples, L = cfg->barcode_len;
    if (S < 1 || S > SGDX_MAX_SAMPLES) return fail(SGDX_EINVAL, "num_samples %u outside 1..%u", S, SGDX_MAX_SAMPLES);
    if (L < 1 || L > SGDX_MAX_BARCODE_LEN)
        return fail(SGDX_EINVAL, "barcode_len %u outside 1..%u", L, SGDX_MAX_BARCODE_LEN);
    if (cfg->matcher > SGDX_MATCHER_PRECOMPUTE) return fail(SGDX_EINVAL, "unknown matcher kind %u", cfg->matcher);
    if (cfg->kernel > SGDX_KERNEL_SEEDED) return fail(SGDX_EINVAL, "unknown kernel kind %u", cfg->kernel);
    if ((cfg->index1_len == 0) != (cfg->index2_len == 0) ||
        (cfg->index1_len && cfg->index1_len + cfg->index2_len != L))
        return fail(SGDX_EINVAL, "index1_len + index2_len must equal barcode_len (or both be 0)");
    if (cfg->index1_len > SGDX_MAX_PACKED_LEN || cfg->index2_len > SGDX_MAX_PACKED_LEN)
        return fail(SGDX_EINVAL, "index halves of %u + %u bases: the hop checker takes at most %u bases per half", cfg->index1_len,
                    cfg->index2_len, SGDX_MAX_PACKED_LEN);
    for (uint64_t i = 0; i < (uint64_t)S * L; i++) {
        uint8_t c = barcodes[i];
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T')
            return fail(SGDX_EINVAL, "expected barcode %llu has byte 0x%02x at %llu (must be upper-case ACGT)",
                        (unsigned long long)(i / L), c, (unsigned long long)(i % L));
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(SGDX_ECUDA, "no CUDA device available (%s); sgdx has no CPU fallback", cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SGDX_EINVAL, "device %d not in 0..%d", cfg->device, ndev - 1);
    CU(cudaSetDevice(cfg->device));

    sgdx_handle *h = new sgdx_handle();
    if (int rc = fill_handle(h, cfg, barcodes)) {
        std::string msg = g_err;  // sgdx_destroy makes CUDA calls of its own
        sgdx_destroy(h);
        g_err = msg;
        return rc;
    }
    *out = h;
    return SGDX_OK;
    });
}

extern "C" void sgdx_destroy(sgdx_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    for (Slot &s : h->slots) {
        if (s.own) cudaStreamSynchronize(s.own);
        cudaFree(s.d_in);
        cudaFree(s.d_idx);
        cudaFree(s.d_dist);
        cudaFree(s.d_extract);
        cudaFree(s.d_route);
        cudaFree(s.d_cexp);
        cudaFree(s.d_dense);
        cudaFree(s.d_exc_idx);
        cudaFree(s.d_exc_rec);
        if (s.h_unm_stats) cudaFreeHost(s.h_unm_stats);
        if (s.e0) cudaEventDestroy(s.e0);
        if (s.e1) cudaEventDestroy(s.e1);
        if (s.own) cudaStreamDestroy(s.own);
    }
    cudaFree(h->d_table);
    cudaFree(h->d_table_long);
    cudaFree(h->d_seed_hdr);
    cudaFree(h->d_seed_ids);
    cudaFree(h->d_ex_slots);
    cudaFree(h->d_ex_ascii);
    cudaFree(h->d_nb_slots);
    cudaFree(h->d_ph_disp);
    cudaFree(h->d_ph_slots);
    cudaFree(h->d_ph_b_disp);
    cudaFree(h->d_ph_b_slots);
    cudaFree(h->d_ph_ms_hdr);
    cudaFree(h->d_ph_ms_ids);
    cudaFree(h->d_ph_nn);
    cudaFree(h->d_ph_hop_slots);
    cudaFree(h->d_ph_hop_disp);
    cudaFree(h->d_hash1);
    cudaFree(h->d_hash2);
    cudaFree(h->d_hop_lut1);
    cudaFree(h->d_hop_lut2);
    cudaFree(h->d_hop_fwd);
    cudaFree(h->d_hop_rev);
    cudaFree(h->d_hop_list);
    cudaFree(h->d_hop_stat);
    cudaFree(h->d_counters);
    post_destroy(h);
    delete h;
}

extern "C" int sgdx_get_config(const sgdx_handle *h, sgdx_config *effective) {
    return sgdx::guarded("sgdx_get_config", [&]() -> int {
    if (!h || !effective) return fail(SGDX_EINVAL, "sgdx_get_config: null argument");
    *effective = h->cfg;
    return SGDX_OK;
    });
}

extern "C" int sgdx_alloc_host(size_t bytes, void **out) {
    return sgdx::guarded("sgdx_alloc_host", [&]() -> int {
    if (!out) return fail(SGDX_EINVAL, "sgdx_alloc_host: null out");
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return SGDX_OK;
    });
}
extern "C" int sgdx_free_host(void *p) {
    return sgdx::guarded("sgdx_free_host", [&]() -> int {
    if (p) CU(cudaFreeHost(p));
    return SGDX_OK;
    });
}

static int check_slot(sgdx_handle *h, uint32_t slot) {
    if (!h) return fail(SGDX_EINVAL, "null handle");
    if (slot >= h->slots.size()) return fail(SGDX_EINVAL, "slot %u >= %zu", slot, h->slots.size());
    return SGDX_OK;
}

static int run_kernels(sgdx_handle *h, Slot &s, const uint8_t *d_ascii, const uint4 *d_packed, uint64_t n,
                       uint16_t *d_idx, uint8_t *d_dist) {
    DevParams p = h->base;
    p.reads_ascii = d_ascii;
    p.reads_packed = d_packed;
    p.n = n;
    p.out_idx = d_idx;
    p.out_dist = d_dist;
    if (h->d_table_long) {  // 33..64 bases
        if (!d_ascii && n) return fail(SGDX_EINVAL, "barcode_len %u: records hold at most %u bases, use the ASCII entry points", p.L, SGDX_MAX_PACKED_LEN);
        CU(cudaEventRecord(s.e0, s.stream));
        CU(launch_long(p, h->d_table_long, h->mode, h->sms, s.stream));
        if (n) h->launches.fetch_add(1);
        CU(cudaEventRecord(s.e1, s.stream));
        s.timed = true;
        return SGDX_OK;
    }
    std::unique_lock<std::mutex> lk(h->post_mu, std::defer_lock);
    if (h->unm.on) lk.lock();  // a downsize of the unmatched table must not overlap a batch
    CU(cudaEventRecord(s.e0, s.stream));
    if (d_packed && h->ph.slots && h->ph.wide) {  // sgdx_match_phx on 16-byte records
        PhParams c = h->ph;
        c.reads_wide = d_packed;
        CU(launch_phx(p, c, h->mode, h->sms, s.stream));
    } else if (h->seeded) {
        CU(launch_seeded(p, h->mode, d_packed != nullptr, h->sms, s.stream));
    } else {
        CU(launch_direct(p, h->mode, d_packed != nullptr, h->sms, s.stream));
    }
    if (n) h->launches.fetch_add(1);
    if (h->unm.on && n)
        if (int rc = post_after_match(h, s, d_ascii, nullptr, d_ascii ? nullptr : d_packed, d_idx, n)) return rc;
    CU(cudaEventRecord(s.e1, s.stream));
    s.timed = true;
    return SGDX_OK;
}

extern "C" int sgdx_match_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_ascii, uint64_t n, uint16_t *d_idx,
                                 uint8_t *d_dist) {
    return sgdx::guarded("sgdx_match_device", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!d_ascii || !d_idx)) return fail(SGDX_EINVAL, "sgdx_match_device: null buffer");
    if (n >> 40) return fail(SGDX_EINVAL, "batch too large");
    CU(cudaSetDevice(h->cfg.device));
    return run_kernels(h, h->slots[slot], d_ascii, nullptr, n, d_idx, d_dist);
    });
}

extern "C" int sgdx_match_packed_device(sgdx_handle *h, uint32_t slot, const sgdx_read *d_packed, uint64_t n,
                                        uint16_t *d_idx, uint8_t *d_dist) {
    return sgdx::guarded("sgdx_match_packed_device", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!d_packed || !d_idx)) return fail(SGDX_EINVAL, "sgdx_match_packed_device: null buffer");
    if (reinterpret_cast<uintptr_t>(d_packed) & 15u) return fail(SGDX_EINVAL, "packed records must be 16-byte aligned");
    CU(cudaSetDevice(h->cfg.device));
    return run_kernels(h, h->slots[slot], nullptr, reinterpret_cast<const uint4 *>(d_packed), n, d_idx, d_dist);
    });
}

extern "C" int sgdx_pack_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_ascii, uint64_t n, sgdx_read *d_packed) {
    return sgdx::guarded("sgdx_pack_device", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!d_ascii || !d_packed)) return fail(SGDX_EINVAL, "sgdx_pack_device: null buffer");
    if (reinterpret_cast<uintptr_t>(d_packed) & 15u) return fail(SGDX_EINVAL, "packed records must be 16-byte aligned");
    if (h->cfg.barcode_len > SGDX_MAX_PACKED_LEN) return fail(SGDX_EINVAL, "sgdx_read records hold at most %u bases", SGDX_MAX_PACKED_LEN);
    CU(cudaSetDevice(h->cfg.device));
    CU(launch_pack(d_ascii, n, h->cfg.barcode_len, reinterpret_cast<uint4 *>(d_packed), h->sms, h->slots[slot].stream));
    if (n) h->launches.fetch_add(1);
    return SGDX_OK;
    });
}

static int ensure_capacity(Slot &s, uint64_t n, uint32_t L) {
    if (n <= s.cap) return SGDX_OK;
    CU(cudaStreamSynchronize(s.stream));
    cudaFree(s.d_in);
    cudaFree(s.d_idx);
    cudaFree(s.d_dist);
    s.d_in = nullptr;
    s.d_idx = nullptr;
    s.d_dist = nullptr;
    s.cap = 0;
    uint64_t cap = n + n / 8 + 1024;
    CU(cudaMalloc(&s.d_in, cap * std::max<uint32_t>(L, 8u)));  // L ASCII bytes or one 64-bit compact record per read
    CU(cudaMalloc(&s.d_idx, cap * sizeof(uint16_t)));
    CU(cudaMalloc(&s.d_dist, cap));
    s.cap = cap;
    return SGDX_OK;
}

extern "C" int sgdx_match_batch(sgdx_handle *h, uint32_t slot, const uint8_t *ascii, uint64_t n, uint16_t *out_idx,
                                uint8_t *out_dist) {
    return sgdx::guarded("sgdx_match_batch", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!ascii || !out_idx)) return fail(SGDX_EINVAL, "sgdx_match_batch: null buffer");
    if (n >> 34) return fail(SGDX_EINVAL, "batch of %llu reads too large; split it", (unsigned long long)n);
    CU(cudaSetDevice(h->cfg.device));
    Slot &s = h->slots[slot];
    const uint32_t L = h->cfg.barcode_len;
    if (int rc = ensure_capacity(s, n, L)) return rc;
    if (n) CU(cudaMemcpyAsync(s.d_in, ascii, n * L, cudaMemcpyHostToDevice, s.stream));
    if (int rc = run_kernels(h, s, s.d_in, nullptr, n, s.d_idx, s.d_dist)) return rc;
    if (n) {
        CU(cudaMemcpyAsync(out_idx, s.d_idx, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, s.stream));
        if (out_dist) CU(cudaMemcpyAsync(out_dist, s.d_dist, n, cudaMemcpyDeviceToHost, s.stream));
    }
    return SGDX_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// compact records (one 64-bit word per read, barcode_len <= 21)
// ------------------------------------------------------------------------------------------------
// the slot's buffer of 16-byte sgdx_read records (expanded compact records, or a packed batch from the host)
static int ensure_packed_scratch(Slot &s, uint64_t n) {
    if (n <= s.cexp_cap) return SGDX_OK;
    CU(cudaStreamSynchronize(s.stream));
    cudaFree(s.d_cexp);
    s.d_cexp = nullptr;
    s.cexp_cap = 0;
    const uint64_t cap = n + n / 8 + 1024;
    CU(cudaMalloc(&s.d_cexp, cap * sizeof(uint4)));
    s.cexp_cap = cap;
    return SGDX_OK;
}

static int run_compact(sgdx_handle *h, Slot &s, const uint64_t *d_records, uint64_t n, uint16_t *d_idx, uint8_t *d_dist) {
    if (h->cfg.barcode_len > kCompactMaxLen)
        return fail(SGDX_EINVAL, "compact records hold at most %u bases, barcode_len is %u", kCompactMaxLen, h->cfg.barcode_len);
    if (reinterpret_cast<uintptr_t>(d_records) & 7u) return fail(SGDX_EINVAL, "compact records must be 8-byte aligned");
    if (h->single || (h->ph.slots && !h->ph.wide)) {
        DevParams p = h->base;
        p.n = n;
        p.out_idx = d_idx;
        p.out_dist = d_dist;
        std::unique_lock<std::mutex> lk(h->post_mu, std::defer_lock);
        if (h->unm.on) lk.lock();  // a downsize of the unmatched table must not overlap a batch
        CU(cudaEventRecord(s.e0, s.stream));
        if (h->single) {  // one barcode: sgdx_match_single
            CU(launch_single(p, d_records, h->mode, h->sms, s.stream));
        } else {  // perfect hash of the match neighbourhood: sgdx_match_ph / sgdx_match_phx
            PhParams c = h->ph;
            c.reads = d_records;
            if (c.radius >= p.M) CU(launch_ph(p, c, h->mode, h->sms, s.stream));
            else CU(launch_phx(p, c, h->mode, h->sms, s.stream));
        }
        if (n) h->launches.fetch_add(1);
        if (h->unm.on && n)  // most frequent unmatched barcodes (demux.rs:553-555), keyed from the records
            if (int rc = post_after_match(h, s, nullptr, d_records, nullptr, d_idx, n)) return rc;
        CU(cudaEventRecord(s.e1, s.stream));
        s.timed = true;
        return SGDX_OK;
    }
    // no perfect hash for this table (too many strings within max_mismatches): 16-byte records, general kernels
    if (int rc = ensure_packed_scratch(s, n)) return rc;
    CU(launch_expand_compact(d_records, n, h->cfg.barcode_len, s.d_cexp, h->sms, s.stream));
    if (n) h->launches.fetch_add(1);
    return run_kernels(h, s, nullptr, s.d_cexp, n, d_idx, d_dist);
}

extern "C" int sgdx_match_compact_device(sgdx_handle *h, uint32_t slot, const uint64_t *d_records, uint64_t n,
                                         uint16_t *d_idx, uint8_t *d_dist) {
    return sgdx::guarded("sgdx_match_compact_device", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!d_records || !d_idx)) return fail(SGDX_EINVAL, "sgdx_match_compact_device: null buffer");
    if (n >> 40) return fail(SGDX_EINVAL, "batch too large");
    CU(cudaSetDevice(h->cfg.device));
    return run_compact(h, h->slots[slot], d_records, n, d_idx, d_dist);
    });
}

extern "C" int sgdx_pack_compact_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_ascii, uint64_t n,
                                        uint64_t *d_records) {
    return sgdx::guarded("sgdx_pack_compact_device", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!d_ascii || !d_records)) return fail(SGDX_EINVAL, "sgdx_pack_compact_device: null buffer");
    if (h->cfg.barcode_len > kCompactMaxLen)
        return fail(SGDX_EINVAL, "compact records hold at most %u bases, barcode_len is %u", kCompactMaxLen, h->cfg.barcode_len);
    if (reinterpret_cast<uintptr_t>(d_records) & 7u) return fail(SGDX_EINVAL, "compact records must be 8-byte aligned");
    CU(cudaSetDevice(h->cfg.device));
    CU(launch_pack_compact(d_ascii, n, h->cfg.barcode_len, d_records, h->sms, h->slots[slot].stream));
    if (n) h->launches.fetch_add(1);
    return SGDX_OK;
    });
}

extern "C" int sgdx_match_compact_batch(sgdx_handle *h, uint32_t slot, const uint64_t *records, uint64_t n,
                                        uint16_t *out_idx, uint8_t *out_dist) {
    return sgdx::guarded("sgdx_match_compact_batch", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!records || !out_idx)) return fail(SGDX_EINVAL, "sgdx_match_compact_batch: null buffer");
    if (n >> 34) return fail(SGDX_EINVAL, "batch of %llu reads too large; split it", (unsigned long long)n);
    CU(cudaSetDevice(h->cfg.device));
    Slot &s = h->slots[slot];
    if (int rc = ensure_capacity(s, n, h->cfg.barcode_len)) return rc;
    uint64_t *d_rec = reinterpret_cast<uint64_t *>(s.d_in);
    if (n) CU(cudaMemcpyAsync(d_rec, records, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
    if (int rc = run_compact(h, s, d_rec, n, s.d_idx, s.d_dist)) return rc;
    if (n) {
        CU(cudaMemcpyAsync(out_idx, s.d_idx, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, s.stream));
        if (out_dist) CU(cudaMemcpyAsync(out_dist, s.d_dist, n, cudaMemcpyDeviceToHost, s.stream));
    }
    return SGDX_OK;
    });
}

// sgdx_match_batch for a host batch of dense records + exception list (include/sgdx.h): 2 L bits per read over PCIe
extern "C" int sgdx_match_dense_batch(sgdx_handle *h, uint32_t slot, const uint8_t *dense, uint64_t n, const uint32_t *exc_index,
                                      const void *exc_record, uint64_t n_exc, uint16_t *out_idx, uint8_t *out_dist) {
    return sgdx::guarded("sgdx_match_dense_batch", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    const uint32_t L = h->cfg.barcode_len;
    if (L > SGDX_MAX_PACKED_LEN) return fail(SGDX_EINVAL, "dense records hold at most %u bases, barcode_len is %u", SGDX_MAX_PACKED_LEN, L);
    if (n && (!dense || !out_idx)) return fail(SGDX_EINVAL, "sgdx_match_dense_batch: null buffer");
    if (n_exc && (!exc_index || !exc_record)) return fail(SGDX_EINVAL, "sgdx_match_dense_batch: null exception list");
    if (n >> 32 || n_exc > n) return fail(SGDX_EINVAL, "sgdx_match_dense_batch: %llu reads / %llu exceptions", (unsigned long long)n, (unsigned long long)n_exc);
    CU(cudaSetDevice(h->cfg.device));
    Slot &s = h->slots[slot];
    const bool wide = L > kCompactMaxLen;  // 16-byte records on the device, 16-byte exception records
    const size_t esz = SGDX_DENSE_EXC_BYTES(L);
    if (int rc = ensure_capacity(s, n, L)) return rc;
    if (wide)
        if (int rc = ensure_packed_scratch(s, n)) return rc;
    const uint64_t bytes = n * SGDX_DENSE_BYTES(L);
    if (bytes + 64 > s.dense_cap || n_exc > s.exc_cap) {
        CU(cudaStreamSynchronize(s.stream));
        if (bytes + 64 > s.dense_cap) {
            cudaFree(s.d_dense);
            s.d_dense = nullptr;
            s.dense_cap = 0;
            const uint64_t cap = bytes + bytes / 8 + 4096;
            CU(cudaMalloc(&s.d_dense, cap));
            s.dense_cap = cap;
        }
        if (n_exc > s.exc_cap) {
            cudaFree(s.d_exc_idx);
            cudaFree(s.d_exc_rec);
            s.d_exc_idx = nullptr;
            s.d_exc_rec = nullptr;
            s.exc_cap = 0;
            const uint64_t cap = n_exc + n_exc / 4 + 1024;
            CU(cudaMalloc(&s.d_exc_idx, cap * sizeof(uint32_t)));
            CU(cudaMalloc(&s.d_exc_rec, cap * 16u));
            s.exc_cap = cap;
        }
    }
    uint64_t *d_rec = reinterpret_cast<uint64_t *>(s.d_in);
    if (n) {
        CU(cudaMemcpyAsync(s.d_dense, dense, bytes, cudaMemcpyHostToDevice, s.stream));
        if (n_exc) {
            CU(cudaMemcpyAsync(s.d_exc_idx, exc_index, n_exc * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
            CU(cudaMemcpyAsync(s.d_exc_rec, exc_record, n_exc * esz, cudaMemcpyHostToDevice, s.stream));
        }
        if (wide)
            CU(launch_unpack_dense_wide(s.d_dense, n, L, s.d_exc_idx, static_cast<const uint4 *>(s.d_exc_rec), n_exc, s.d_cexp, h->sms, s.stream));
        else
            CU(launch_unpack_dense(s.d_dense, n, L, s.d_exc_idx, static_cast<const uint64_t *>(s.d_exc_rec), n_exc, d_rec, h->sms, s.stream));
        h->launches.fetch_add(n_exc ? 2 : 1);
    }
    if (wide) {
        if (int rc = run_kernels(h, s, nullptr, s.d_cexp, n, s.d_idx, s.d_dist)) return rc;
    } else {
        if (int rc = run_compact(h, s, d_rec, n, s.d_idx, s.d_dist)) return rc;
    }
    if (n) {
        CU(cudaMemcpyAsync(out_idx, s.d_idx, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, s.stream));
        if (out_dist) CU(cudaMemcpyAsync(out_dist, s.d_dist, n, cudaMemcpyDeviceToHost, s.stream));
    }
    return SGDX_OK;
    });
}

// sgdx_match_batch for a host buffer of 16-byte sgdx_read records (any barcode_len <= 32; sgdx_pack_host writes them)
extern "C" int sgdx_match_packed_batch(sgdx_handle *h, uint32_t slot, const sgdx_read *records, uint64_t n,
                                       uint16_t *out_idx, uint8_t *out_dist) {
    return sgdx::guarded("sgdx_match_packed_batch", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    if (n && (!records || !out_idx)) return fail(SGDX_EINVAL, "sgdx_match_packed_batch: null buffer");
    if (n >> 34) return fail(SGDX_EINVAL, "batch of %llu reads too large; split it", (unsigned long long)n);
    CU(cudaSetDevice(h->cfg.device));
    Slot &s = h->slots[slot];
    if (int rc = ensure_capacity(s, n, h->cfg.barcode_len)) return rc;
    if (int rc = ensure_packed_scratch(s, n)) return rc;
    if (n) CU(cudaMemcpyAsync(s.d_cexp, records, n * sizeof(uint4), cudaMemcpyHostToDevice, s.stream));
    if (int rc = run_kernels(h, s, nullptr, s.d_cexp, n, s.d_idx, s.d_dist)) return rc;
    if (n) {
        CU(cudaMemcpyAsync(out_idx, s.d_idx, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, s.stream));
        if (out_dist) CU(cudaMemcpyAsync(out_dist, s.d_dist, n, cudaMemcpyDeviceToHost, s.stream));
    }
    return SGDX_OK;
    });
}

extern "C" int sgdx_sync(sgdx_handle *h, uint32_t slot, float *device_ms) {
    return sgdx::guarded("sgdx_sync", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    Slot &s = h->slots[slot];
    CU(cudaStreamSynchronize(s.stream));
    if (device_ms) {
        *device_ms = 0.f;
        if (s.timed) CU(cudaEventElapsedTime(device_ms, s.e0, s.e1));
    }
    return post_on_sync(h, s);
    });
}

extern "C" int sgdx_set_stream(sgdx_handle *h, uint32_t slot, void *stream) {
    return sgdx::guarded("sgdx_set_stream", [&]() -> int {
    if (int rc = check_slot(h, slot)) return rc;
    Slot &s = h->slots[slot];
    CU(cudaStreamSynchronize(s.stream));
    s.stream = stream ? static_cast<cudaStream_t>(stream) : s.own;
    s.timed = false;
    return SGDX_OK;
    });
}

extern "C" const char *sgdx_kernel_name(const sgdx_handle *h) {
    if (!h) return "";
    if (h->d_table_long) return "sgdx_match_long";
    if (!h->seeded) return "sgdx_match_direct";
    return h->base.ex_on ? "sgdx_match_exact_first" : "sgdx_match_seeded";
}

extern "C" const char *sgdx_compact_kernel_name(const sgdx_handle *h) {
    if (!h || h->cfg.barcode_len > kCompactMaxLen) return "";
    if (h->single) return "sgdx_match_single";
    if (h->ph.slots && !h->ph.wide) return h->ph.radius >= h->base.M ? "sgdx_match_ph" : "sgdx_match_phx";
    return sgdx_kernel_name(h);  // no perfect hash: expanded records, general kernels
}

extern "C" const char *sgdx_packed_kernel_name(const sgdx_handle *h) {
    if (!h || h->d_table_long) return "";
    if (h->ph.slots && h->ph.wide) return "sgdx_match_phx";
    return h->seeded ? "sgdx_match_seeded" : "sgdx_match_direct";
}

extern "C" uint32_t sgdx_path_flags(const sgdx_handle *h) {
    if (!h) return 0u;
    const DevParams &p = h->base;
    return (h->seeded ? SGDX_PATH_SEED_INDEX : 0u) | (h->seeded && p.ex_on ? SGDX_PATH_EXACT : 0u) |
           (h->seeded && p.ex_on && p.nb_on ? SGDX_PATH_NEIGHBOURS : 0u) | (p.hop_direct ? SGDX_PATH_HOP_DIRECT : 0u) |
           (h->ph.slots ? SGDX_PATH_PERFECT_HASH : 0u) | (h->ph.slots && h->ph.ms_C ? SGDX_PATH_SHORT_SEED : 0u);
}

extern "C" uint64_t sgdx_launch_count(const sgdx_handle *h) { return h ? h->launches.load() : 0; }

// The hop counters are two dense u1 x u2 matrices (38 MB at 1536 unique dual indexes, 1 GB at the sample limit) of
// which real runs touch a few thousand cells: the non-zero ones are compacted on the device and only those cross PCIe.
// entry = {cell (+ cells for the reversed-orientation matrix), count}; stat[0] = entries, stat[1] = sum of the counts
__global__ void __launch_bounds__(256) sgdx_hop_compact(const unsigned long long *__restrict__ fwd,
                                                        const unsigned long long *__restrict__ rev, uint64_t cells,
                                                        ulonglong2 *__restrict__ out, uint64_t cap, unsigned long long *stat) {
    unsigned long long local = 0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * cells; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = i < cells ? fwd[i] : rev[i - cells];
        if (c) {
            local += c;
            const unsigned long long at = atomicAdd(stat, 1ull);
            if (at < cap) out[at] = make_ulonglong2(i, c);
        }
    }
    for (int d = 16; d; d >>= 1) local += __shfl_xor_sync(0xFFFFFFFFu, local, d);
    if ((threadIdx.x & 31u) == 0u && local) atomicAdd(stat + 1, local);
}

// want_list = false: only the sum of the counts.  Device idle (callers synchronise first).
static int fetch_hop(sgdx_handle *h, bool want_list, std::vector<ulonglong2> &list, uint64_t *hop_reads) {
    const uint64_t cells = (uint64_t)h->base.u1 * h->base.u2;
    std::lock_guard<std::mutex> lk(h->hop_mu);
    if (!h->d_hop_stat) CU(cudaMalloc(&h->d_hop_stat, 2 * sizeof(unsigned long long)));
    if (want_list && !h->d_hop_list) {
        h->hop_list_cap = 1u << 16;
        CU(cudaMalloc(&h->d_hop_list, h->hop_list_cap * sizeof(ulonglong2)));
    }
    const uint32_t grid = (uint32_t)std::min<uint64_t>((2 * cells + 255) / 256, (uint64_t)h->sms * 8u);
    for (int attempt = 0; attempt < 2; attempt++) {
        unsigned long long st[2] = {0ull, 0ull};
        CU(cudaMemset(h->d_hop_stat, 0, sizeof st));
        sgdx_hop_compact<<<grid, 256>>>(h->d_hop_fwd, h->d_hop_rev, cells, h->d_hop_list, want_list ? h->hop_list_cap : 0, h->d_hop_stat);
        CU(cudaGetLastError());
        CU(cudaMemcpy(st, h->d_hop_stat, sizeof st, cudaMemcpyDeviceToHost));
        if (hop_reads) *hop_reads = st[1];
        if (!want_list) return SGDX_OK;
        if (st[0] <= h->hop_list_cap) {
            list.resize(st[0]);
            if (st[0]) CU(cudaMemcpy(list.data(), h->d_hop_list, st[0] * sizeof(ulonglong2), cudaMemcpyDeviceToHost));
            return SGDX_OK;
        }
        cudaFree(h->d_hop_list);  // more non-zero cells than the list holds: once more with room for all of them
        h->d_hop_list = nullptr;
        h->hop_list_cap = st[0] + st[0] / 8;
        CU(cudaMalloc(&h->d_hop_list, h->hop_list_cap * sizeof(ulonglong2)));
    }
    return fail(SGDX_ECUDA, "hop counters changed while they were read");
}

extern "C" int sgdx_counters(sgdx_handle *h, uint64_t *per_sample, sgdx_totals *totals) {
    return sgdx::guarded("sgdx_counters", [&]() -> int {
    if (!h) return fail(SGDX_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    const uint32_t S = h->cfg.num_samples;
    std::vector<unsigned long long> raw((size_t)(S + 1) * 3);
    CU(cudaMemcpy(raw.data(), h->d_counters, raw.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    sgdx_totals t{};
    for (uint32_t b = 0; b <= S; b++) {
        uint64_t other = raw[b * 3 + 0], perfect = raw[b * 3 + 1], one = raw[b * 3 + 2];
        uint64_t total = other + perfect + one;
        if (per_sample) {
            per_sample[b * 3 + 0] = total;
            per_sample[b * 3 + 1] = perfect;
            per_sample[b * 3 + 2] = one;
        }
        t.total_templates += total;
        (b < S ? t.matched : t.undetermined) += total;
    }
    if (h->base.hop_on) {
        std::vector<ulonglong2> none;
        if (int rc = fetch_hop(h, false, none, &t.hop_reads)) return rc;
    }
    if (totals) *totals = t;
    return SGDX_OK;
    });
}

extern "C" int sgdx_reset_counters(sgdx_handle *h) {
    return sgdx::guarded("sgdx_reset_counters", [&]() -> int {
    if (!h) return fail(SGDX_EINVAL, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(h->d_counters, 0, (size_t)(h->cfg.num_samples + 1) * 3 * sizeof(unsigned long long)));
    if (h->base.hop_on) {
        size_t cells = (size_t)h->base.u1 * h->base.u2;
        CU(cudaMemset(h->d_hop_fwd, 0, cells * sizeof(unsigned long long)));
        CU(cudaMemset(h->d_hop_rev, 0, cells * sizeof(unsigned long long)));
    }
    if (int rc = post_reset(h)) return rc;
    CU(cudaDeviceSynchronize());  // the memsets are on the default stream, batches on non-blocking ones
    return SGDX_OK;
    });
}

// dense (index1 id, index2 id) cells -> {observed barcode: count}; a barcode that satisfies the forward
// rule is always booked in hop_fwd, so the two tables never hold the same string twice -- but sum anyway.
static int hop_entries(sgdx_handle *h, std::map<std::string, uint64_t> &out) {
    if (!h->base.hop_on) return SGDX_OK;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    std::vector<ulonglong2> list;
    if (int rc = fetch_hop(h, true, list, nullptr)) return rc;
    const uint64_t u2 = h->base.u2, cells = (uint64_t)h->base.u1 * u2;
    for (const ulonglong2 &e : list) {
        const uint64_t i = e.x < cells ? e.x : e.x - cells;
        if (e.x < cells) out[h->keys1[i / u2] + h->keys2[i % u2]] += e.y;
        else out[h->keys2[i % u2] + h->keys1[i / u2]] += e.y;
    }
    return SGDX_OK;
}

extern "C" int sgdx_hop_count(sgdx_handle *h, uint64_t *n_keys) {
    return sgdx::guarded("sgdx_hop_count", [&]() -> int {
    if (!h || !n_keys) return fail(SGDX_EINVAL, "sgdx_hop_count: null argument");
    std::map<std::string, uint64_t> m;
    if (int rc = hop_entries(h, m)) return rc;
    *n_keys = m.size();
    return SGDX_OK;
    });
}

extern "C" int sgdx_hop_export(sgdx_handle *h, uint8_t *keys, uint64_t *counts, uint64_t cap, uint64_t *n_written) {
    return sgdx::guarded("sgdx_hop_export", [&]() -> int {
    if (!h || !n_written) return fail(SGDX_EINVAL, "sgdx_hop_export: null argument");
    std::map<std::string, uint64_t> m;
    if (int rc = hop_entries(h, m)) return rc;
    if (m.size() > cap) return fail(SGDX_EINVAL, "hop export needs room for %zu keys, got %llu", m.size(), (unsigned long long)cap);
    if (!m.empty() && (!keys || !counts)) return fail(SGDX_EINVAL, "sgdx_hop_export: null output buffer");
    const uint32_t L = h->cfg.barcode_len;
    uint64_t i = 0;
    for (auto &kv : m) {
        memcpy(keys + i * L, kv.first.data(), L);
        counts[i++] = kv.second;
    }
    *n_written = i;
    return SGDX_OK;
    });
}

extern "C" int sgdx_measure_int_peak(int device, double *popc_per_s, double *lop3_per_s) {
    return sgdx::guarded("sgdx_measure_int_peak", [&]() -> int {
    if (!popc_per_s || !lop3_per_s) return fail(SGDX_EINVAL, "sgdx_measure_int_peak: null argument");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || device < 0 || device >= ndev) return fail(SGDX_ECUDA, "no such CUDA device %d", device);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    CU(measure_int_peak(prop.multiProcessorCount, popc_per_s, lop3_per_s));
    return SGDX_OK;
    });
}

// ------------------------------------------------------------------------------------------------
// batching stage
// ------------------------------------------------------------------------------------------------
struct Batch {
    uint16_t *idx = nullptr;  // pinned
    uint8_t *dist = nullptr;  // pinned
    uint64_t n = 0;
    uint32_t slot = 0;
};

struct sgdx_batcher {
    sgdx_handle *h = nullptr;
    uint64_t batch_reads = 0;
    uint32_t L = 0;
    bool compact = false;           // the pinned buffers hold 64-bit compact records instead of L ASCII bytes
    bool dense = false;             // ... or dense records (2 bits per base) + the exception lists below
    std::vector<uint8_t *> in;      // one pinned input buffer per slot
    std::vector<uint32_t *> exc_idx;  // dense mode: pinned exception lists per slot (room for every read of a batch)
    std::vector<uint8_t *> exc_rec;   //   (SGDX_DENSE_EXC_BYTES per entry)
    uint64_t n_exc = 0;             // exceptions of the batch being filled
    std::vector<bool> slot_busy;    // slot has an unsynced batch
    uint32_t cur = 0;               // slot being filled
    uint64_t fill = 0;
    std::deque<Batch> inflight;     // submission order
    std::vector<Batch> free_results;
    Batch handed{};                 // batch whose pointers the caller currently holds
};

extern "C" int sgdx_batcher_create(sgdx_handle *h, uint64_t batch_reads, sgdx_batcher **out) {
    return sgdx::guarded("sgdx_batcher_create", [&]() -> int {
    if (!h || !out || batch_reads == 0) return fail(SGDX_EINVAL, "sgdx_batcher_create: bad argument");
    sgdx_batcher *b = new sgdx_batcher();
    b->h = h;
    b->batch_reads = batch_reads;
    b->L = h->cfg.barcode_len;
    b->in.resize(h->slots.size(), nullptr);
    b->slot_busy.assign(h->slots.size(), false);
    for (auto &p : b->in) {
        void *q = nullptr;
        if (int rc = sgdx_alloc_host(batch_reads * std::max<uint32_t>(b->L, 8u), &q)) {
            sgdx_batcher_destroy(b);
            return rc;
        }
        p = static_cast<uint8_t *>(q);
    }
    *out = b;
    return SGDX_OK;
    });
}

extern "C" int sgdx_batcher_set_compact(sgdx_batcher *b, int on) {
    return sgdx::guarded("sgdx_batcher_set_compact", [&]() -> int {
    if (!b) return fail(SGDX_EINVAL, "null batcher");
    if (b->fill || !b->inflight.empty()) return fail(SGDX_ESTATE, "sgdx_batcher_set_compact: batches are in flight");
    if (on == 1 && b->L > kCompactMaxLen)
        return fail(SGDX_EINVAL, "compact records hold at most %u bases, barcode_len is %u", kCompactMaxLen, b->L);
    if (on == 2 && b->L > SGDX_MAX_PACKED_LEN)
        return fail(SGDX_EINVAL, "dense records hold at most %u bases, barcode_len is %u", SGDX_MAX_PACKED_LEN, b->L);
    if (on == 2 && b->exc_idx.empty()) {  // dense records: the exception lists, worst case every read
        if (b->batch_reads >> 32) return fail(SGDX_EINVAL, "dense records: batches of at most 2^32 - 1 reads");
        b->exc_idx.assign(b->in.size(), nullptr);
        b->exc_rec.assign(b->in.size(), nullptr);
        for (size_t i = 0; i < b->in.size(); i++) {
            void *q = nullptr;
            if (int rc = sgdx_alloc_host(b->batch_reads * sizeof(uint32_t), &q)) return rc;
            b->exc_idx[i] = static_cast<uint32_t *>(q);
            if (int rc = sgdx_alloc_host(b->batch_reads * SGDX_DENSE_EXC_BYTES(b->L), &q)) return rc;
            b->exc_rec[i] = static_cast<uint8_t *>(q);
        }
    }
    b->compact = on == 1;
    b->dense = on == 2;
    return SGDX_OK;
    });
}

static void release_result(sgdx_batcher *b, Batch &r) {
    if (r.idx) b->free_results.push_back(r);
    r = Batch{};
}

extern "C" void sgdx_batcher_destroy(sgdx_batcher *b) {
    if (!b) return;
    for (uint32_t s = 0; s < b->slot_busy.size(); s++)
        if (b->slot_busy[s]) sgdx_sync(b->h, s, nullptr);
    for (auto p : b->in) cudaFreeHost(p);
    for (auto p : b->exc_idx) cudaFreeHost(p);
    for (auto p : b->exc_rec) cudaFreeHost(p);
    release_result(b, b->handed);
    for (auto &r : b->inflight) b->free_results.push_back(r);
    for (auto &r : b->free_results) {
        cudaFreeHost(r.idx);
        cudaFreeHost(r.dist);
    }
    delete b;
}

static int submit(sgdx_batcher *b) {
    if (b->fill == 0) return SGDX_OK;
    if (b->inflight.size() >= 2 * b->in.size())  // every waiting batch holds pinned result buffers
        return fail(SGDX_ESTATE, "sgdx_batcher: %zu batches wait for sgdx_batcher_next", b->inflight.size());
    Batch r;
    if (!b->free_results.empty()) {
        r = b->free_results.back();
        b->free_results.pop_back();
    } else {
        void *q = nullptr;
        if (int rc = sgdx_alloc_host(b->batch_reads * sizeof(uint16_t), &q)) return rc;
        r.idx = static_cast<uint16_t *>(q);
        if (int rc = sgdx_alloc_host(b->batch_reads, &q)) {
            cudaFreeHost(r.idx);
            return rc;
        }
        r.dist = static_cast<uint8_t *>(q);
    }
    r.n = b->fill;
    r.slot = b->cur;
    const int rc_submit =
        b->dense ? sgdx_match_dense_batch(b->h, b->cur, b->in[b->cur], r.n, b->exc_idx[b->cur], b->exc_rec[b->cur], b->n_exc, r.idx, r.dist)
        : b->compact ? sgdx_match_compact_batch(b->h, b->cur, reinterpret_cast<const uint64_t *>(b->in[b->cur]), r.n, r.idx, r.dist)
                     : sgdx_match_batch(b->h, b->cur, b->in[b->cur], r.n, r.idx, r.dist);
    if (int rc = rc_submit) {
        b->free_results.push_back(r);  // the reads stay in the slot's input buffer; a retry re-submits them
        return rc;
    }
    b->slot_busy[b->cur] = true;
    b->inflight.push_back(r);
    b->fill = 0;
    b->n_exc = 0;
    b->cur = (b->cur + 1) % (uint32_t)b->in.size();
    if (b->slot_busy[b->cur]) {  // about to overwrite this slot's input buffer: its batch must have left the host
        if (int rc = sgdx_sync(b->h, b->cur, nullptr)) return rc;
        b->slot_busy[b->cur] = false;
    }
    return SGDX_OK;
}

extern "C" int sgdx_batcher_push(sgdx_batcher *b, const uint8_t *ascii, uint64_t n) {
    return sgdx::guarded("sgdx_batcher_push", [&]() -> int {
    return sgdx_batcher_push_n(b, ascii, n, nullptr);
    });
}

extern "C" int sgdx_batcher_push_n(sgdx_batcher *b, const uint8_t *ascii, uint64_t n, uint64_t *consumed) {
    return sgdx::guarded("sgdx_batcher_push_n", [&]() -> int {
    if (consumed) *consumed = 0;
    if (!b || (n && !ascii)) return fail(SGDX_EINVAL, "sgdx_batcher_push: null argument");
    while (n) {
        // a full buffer that could not be submitted earlier (failed launch, results not taken) goes first
        if (b->fill == b->batch_reads)
            if (int rc = submit(b)) return rc;
        uint64_t take = std::min(n, b->batch_reads - b->fill);
        if (b->dense) {  // the copy into the pinned buffer is the packing pass: 2 bits per base + exceptions
            uint64_t k = 0;
            uint32_t *ei = b->exc_idx[b->cur] + b->n_exc;
            if (int rc = sgdx_pack_dense_host(ascii, take, b->L, b->in[b->cur] + b->fill * SGDX_DENSE_BYTES(b->L), ei,
                                              b->exc_rec[b->cur] + b->n_exc * SGDX_DENSE_EXC_BYTES(b->L), b->batch_reads - b->n_exc, &k, 1))
                return rc;
            for (uint64_t e = 0; e < k; e++) ei[e] += (uint32_t)b->fill;  // read numbers of the batch, not of this push
            b->n_exc += k;
        } else if (b->compact) {  // the copy into the pinned buffer is the packing pass
            if (int rc = sgdx_pack_compact_host(ascii, take, b->L, reinterpret_cast<uint64_t *>(b->in[b->cur]) + b->fill, 1))
                return rc;
        } else {
            memcpy(b->in[b->cur] + b->fill * b->L, ascii, take * b->L);
        }
        b->fill += take;
        ascii += take * b->L;
        n -= take;
        if (consumed) *consumed += take;
        if (b->fill == b->batch_reads)
            if (int rc = submit(b)) return rc;
    }
    return SGDX_OK;
    });
}

extern "C" int sgdx_batcher_flush(sgdx_batcher *b) {
    return sgdx::guarded("sgdx_batcher_flush", [&]() -> int {
    if (!b) return fail(SGDX_EINVAL, "null batcher");
    return submit(b);
    });
}

extern "C" int sgdx_batcher_next(sgdx_batcher *b, const uint16_t **idx, const uint8_t **dist, uint64_t *n) {
    return sgdx::guarded("sgdx_batcher_next", [&]() -> int {
    if (!b || !idx || !dist || !n) return fail(SGDX_EINVAL, "sgdx_batcher_next: null argument");
    release_result(b, b->handed);
    *idx = nullptr;
    *dist = nullptr;
    *n = 0;
    if (b->inflight.empty()) return SGDX_OK;
    Batch r = b->inflight.front();
    b->inflight.pop_front();
    // a slot's stream is in-order, so syncing it covers this batch even if a newer one was queued on it
    if (int rc = sgdx_sync(b->h, r.slot, nullptr)) return rc;
    bool newer_on_slot = false;
    for (auto &q : b->inflight) newer_on_slot |= (q.slot == r.slot);
    if (!newer_on_slot) b->slot_busy[r.slot] = false;
    b->handed = r;
    *idx = r.idx;
    *dist = r.dist;
    *n = r.n;
    return SGDX_OK;
    });
}
