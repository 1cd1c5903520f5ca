// sgdx_synth_core.h -- the seeded synthetic generator of SURVEY Appendix C as plain C++ (host + device).
// Included by sgdx_synth.cu (libsgdx.so: host and CUDA generators) and by oracle/synth_host.cpp, so that the CPU
// reference arm of bench.py can draw the same reads without loading the product library.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sgdx_synth.h"

#ifdef __CUDACC__
#define SGDX_HD __host__ __device__
#else
#define SGDX_HD
#endif

namespace sgdx {

SGDX_HD inline uint64_t synth_mix64(uint64_t z) {  // splitmix64 finaliser (same function as mix64 of sgdx_device.cuh)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Rng {
    uint64_t s;
    SGDX_HD explicit Rng(uint64_t seed, uint64_t id) : s(synth_mix64(seed ^ synth_mix64(id + 0x632BE59BD9B4E019ull))) {}
    SGDX_HD uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        return synth_mix64(s);
    }
};

SGDX_HD inline uint32_t pick_sample(Rng &g, uint32_t S, uint32_t zipf) {
    uint64_t v = g.next() >> 32;
    if (zipf) v = (v * v) >> 32;  // density ~ 1/sqrt(k): a few popular samples, a long tail
    return (uint32_t)((v * S) >> 32);
}

SGDX_HD inline void synth_one(const sgdx_synth_params &p, const uint8_t *table, uint64_t id, uint8_t *out) {
    const char bases[4] = {'A', 'C', 'G', 'T'};
    const uint32_t L = p.barcode_len, S = p.num_samples;
    Rng g(p.seed, id);
    uint32_t kind = (uint32_t)(g.next() >> 32);
    bool templated = true;
    uint64_t t_hop = (uint64_t)p.p_true + (p.index1_len ? p.p_hop : 0u);
    if (kind < p.p_true) {
        uint32_t k = pick_sample(g, S, p.zipf);
        for (uint32_t j = 0; j < L; j++) out[j] = table[(uint64_t)k * L + j];
    } else if ((uint64_t)kind < t_hop) {
        uint32_t a = pick_sample(g, S, p.zipf), b = pick_sample(g, S, p.zipf);
        if (b == a) b = (a + 1) % S;
        for (uint32_t j = 0; j < L; j++) out[j] = table[(uint64_t)(j < p.index1_len ? a : b) * L + j];
    } else {
        templated = false;
        uint64_t bits = g.next();
        for (uint32_t j = 0; j < L; j++) out[j] = (uint8_t)bases[(bits >> (2 * j)) & 3u];
    }
    for (uint32_t j = 0; j < L; j++) {
        uint64_t u = g.next();
        uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
        if (templated && lo < p.p_sub) {
            uint32_t cur = out[j] == 'A' ? 0u : out[j] == 'C' ? 1u : out[j] == 'G' ? 2u : 3u;
            out[j] = (uint8_t)bases[(cur + 1u + (uint32_t)(g.next() % 3u)) & 3u];
        }
        if (hi < p.p_n) out[j] = 'N';
        else if ((uint64_t)hi < (uint64_t)p.p_n + p.p_x) out[j] = 'R';
    }
}

inline uint32_t synth_hamming(const uint8_t *a, const uint8_t *b, uint32_t n) {
    uint32_t d = 0;
    for (uint32_t i = 0; i < n; i++) d += a[i] != b[i];
    return d;
}

// greedy: accept a random string iff its distance to every accepted one is >= min_dist
inline bool draw_set(uint64_t seed, uint32_t count, uint32_t len, uint32_t min_dist, std::vector<std::string> &out) {
    const char bases[4] = {'A', 'C', 'G', 'T'};
    Rng g(seed, len * 1315423911ull + count);
    out.clear();
    for (uint64_t tries = 0; out.size() < count; tries++) {
        if (tries > 4000000ull + 4000ull * count) return false;
        std::string s(len, 'A');
        for (uint32_t j = 0; j < len; j += 32) {
            uint64_t bits = g.next();
            for (uint32_t k = 0; k < 32 && j + k < len; k++) s[j + k] = bases[(bits >> (2 * k)) & 3u];
        }
        bool ok = true;
        for (auto &t : out)
            if (synth_hamming((const uint8_t *)s.data(), (const uint8_t *)t.data(), len) < min_dist) {
                ok = false;
                break;
            }
        if (ok) out.push_back(s);
    }
    return true;
}


constexpr uint32_t kSynthMaxLen = 32;

inline int synth_table_impl(uint64_t seed, uint32_t S, uint32_t L, uint32_t i1, uint32_t min_dist, uint8_t *out) {
    if (!out || S == 0 || L == 0 || L > kSynthMaxLen || i1 >= L) return -1;
    std::vector<std::string> a, b;
    if (i1 == 0) {
        if (!draw_set(seed, S, L, min_dist, a)) return -1;
        for (uint32_t s = 0; s < S; s++) memcpy(out + (uint64_t)s * L, a[s].data(), L);
        return 0;
    }
    if (!draw_set(seed, S, i1, min_dist, a) || !draw_set(seed ^ 0xA5A5A5A5DEADBEEFull, S, L - i1, min_dist, b)) return -1;
    for (uint32_t s = 0; s < S; s++) {
        memcpy(out + (uint64_t)s * L, a[s].data(), i1);
        memcpy(out + (uint64_t)s * L + i1, b[s].data(), L - i1);
    }
    return 0;
}

inline int synth_reads_host_impl(const sgdx_synth_params *p, const uint8_t *table, uint64_t first, uint64_t n, uint8_t *out,
                                 int threads) {
    if (!p || !table || (n && !out) || p->barcode_len == 0 || p->barcode_len > kSynthMaxLen) return -1;
    if (threads < 1) threads = 1;
    auto work = [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; i++) synth_one(*p, table, first + i, out + i * p->barcode_len);
    };
    if (threads == 1 || n < 4096) {
        work(0, n);
    } else {
        std::vector<std::thread> ts;
        for (int t = 0; t < threads; t++) ts.emplace_back(work, n * t / threads, n * (t + 1) / threads);
        for (auto &t : ts) t.join();
    }
    return 0;
}

}  // namespace sgdx
