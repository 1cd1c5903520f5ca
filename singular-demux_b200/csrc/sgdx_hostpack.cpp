// sgdx_hostpack.cpp -- host side of the compact read format (include/sgdx.h, sgdx_pack_compact_host).
//
// The batching stage between the segment extractor (demux.rs:504-522) and the matcher writes every observed
// barcode into a pinned buffer anyway; writing it as one 64-bit compact record (sgdx_compact.cu: lo plane |
// hi plane << 21 | invalid plane << 42) instead of L ASCII bytes is the same pass over the bytes and cuts the
// PCIe traffic of a 20-base read from 20 to 8 bytes.  Pure host code: this is a data-format conversion, not a
// matcher -- assignment itself has no CPU path.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "sgdx_guard.h"
#include "../../include/sgdx.h"

namespace sgdx {
int fail(int code, const char *fmt, ...);  // sgdx_api.cu: sets the thread's sgdx_last_error text
}

namespace {

// test hook (latched at first use): run the scalar packer on a CPU that has AVX2; ignored with -DSGDX_NO_TEST_HOOKS
inline bool no_avx2_hook() {
#ifdef SGDX_NO_TEST_HOOKS
    return false;
#else
    return getenv("SGDX_PACK_NO_AVX2") != nullptr;
#endif
}

constexpr uint32_t kField = 21;

inline uint64_t pack_scalar(const uint8_t *p, uint32_t L) {
    uint32_t lo = 0, hi = 0, inv = 0;
    for (uint32_t j = 0; j < L; j++) {
        const uint8_t c = p[j];
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') {
            lo |= (uint32_t)((c >> 1) & 1u) << j;
            hi |= (uint32_t)((c >> 2) & 1u) << j;
        } else {
            inv |= 1u << j;
            if (c != 'N') lo |= 1u << j;  // foreign byte: mismatches every sample, is not a no-call
        }
    }
    return (uint64_t)lo | ((uint64_t)hi << kField) | ((uint64_t)inv << (2 * kField));
}

void pack_range_scalar(const uint8_t *ascii, uint64_t begin, uint64_t end, uint32_t L, uint64_t *out) {
    for (uint64_t r = begin; r < end; r++) out[r] = pack_scalar(ascii + r * L, L);
}

#if defined(__x86_64__)
// 32 bytes per load, one read per iteration: bit 1 / bit 2 of every byte through a shift + movemask, validity
// through a nibble-indexed table of the four letters (a byte is A/C/G/T iff table[byte & 15] == byte).
// Measured and not kept: 64-byte AVX-512 loads (three reads per load, masks cut with PEXT) and non-temporal
// stores of the records -- neither was faster; the pass runs at the speed of a memcpy of the same bytes.
__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t *ascii, uint64_t begin, uint64_t end, uint32_t L,
                                                     uint64_t total, uint64_t *out) {
    // reads whose 32-byte load would run past the end of the buffer take the scalar path
    const uint64_t bytes = total * L;
    uint64_t safe_end = end;
    while (safe_end > begin && (safe_end - 1) * L + 32 > bytes) safe_end--;
    // unused nibbles hold 0x01, whose own nibble (1) is taken by 'A': no byte equals its entry by accident
    const __m256i lut = _mm256_setr_epi8(1, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1,
                                         1, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1);
    const __m256i nib = _mm256_set1_epi8(0x0F), enn = _mm256_set1_epi8('N');
    const uint32_t m = (1u << L) - 1u;  // L <= 21
    for (uint64_t r = begin; r < safe_end; r++) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(ascii + r * L));
        const uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 6));
        const uint32_t b2 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 5));
        const uint32_t ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, _mm256_and_si256(x, nib)), x));
        if (__builtin_expect((ok & m) == m, 1)) {  // every base is A/C/G/T: no invalid plane to form
            out[r] = (uint64_t)(b1 & m) | ((uint64_t)(b2 & m) << kField);
            continue;
        }
        const uint32_t isn = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, enn));
        const uint32_t inv = ~ok & m;
        const uint32_t lo = (b1 & ok & m) | (inv & ~isn);
        const uint32_t hi = b2 & ok & m;
        out[r] = (uint64_t)lo | ((uint64_t)hi << kField) | ((uint64_t)inv << (2 * kField));
    }
    pack_range_scalar(ascii, safe_end, end, L, out);
}
#endif

// ---- 16-byte sgdx_read records {lo, hi, nmask, xmask}: any barcode_len <= 32 ---------------------------------
void planes_range_scalar(const uint8_t *ascii, uint64_t begin, uint64_t end, uint32_t L, sgdx_read *out) {
    for (uint64_t r = begin; r < end; r++) {
        const uint8_t *p = ascii + r * L;
        sgdx_read v = {0u, 0u, 0u, 0u};
        for (uint32_t j = 0; j < L; j++) {
            const uint8_t c = p[j];
            if (c == 'A' || c == 'C' || c == 'G' || c == 'T') {
                v.lo |= (uint32_t)((c >> 1) & 1u) << j;
                v.hi |= (uint32_t)((c >> 2) & 1u) << j;
            } else if (c == 'N') {
                v.nmask |= 1u << j;
            } else {
                v.xmask |= 1u << j;
            }
        }
        out[r] = v;
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void planes_range_avx2(const uint8_t *ascii, uint64_t begin, uint64_t end, uint32_t L,
                                                       uint64_t total, sgdx_read *out) {
    const uint64_t bytes = total * L;
    uint64_t safe_end = end;
    while (safe_end > begin && (safe_end - 1) * L + 32 > bytes) safe_end--;
    const __m256i lut = _mm256_setr_epi8(1, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1,
                                         1, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1);
    const __m256i nib = _mm256_set1_epi8(0x0F), enn = _mm256_set1_epi8('N');
    const uint32_t m = L >= 32 ? 0xFFFFFFFFu : (1u << L) - 1u;
    for (uint64_t r = begin; r < safe_end; r++) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(ascii + r * L));
        const uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 6));
        const uint32_t b2 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 5));
        const uint32_t ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, _mm256_and_si256(x, nib)), x));
        const uint32_t isn = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, enn));
        sgdx_read v;
        v.lo = b1 & ok & m;
        v.hi = b2 & ok & m;
        v.nmask = isn & m;
        v.xmask = ~ok & ~isn & m;
        out[r] = v;
    }
    planes_range_scalar(ascii, safe_end, end, L, out);
}
#endif

void planes_range(const uint8_t *ascii, uint64_t begin, uint64_t end, uint32_t L, uint64_t total, sgdx_read *out) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2") && !no_avx2_hook();
    if (have_avx2) return planes_range_avx2(ascii, begin, end, L, total, out);
#endif
    (void)total;
    planes_range_scalar(ascii, begin, end, L, out);
}

template <typename Out, typename Fn>
void run_threads(Fn fn, const uint8_t *ascii, uint64_t n, uint32_t L, Out *out, uint32_t threads) {
    if (threads == 0) threads = 1;
    const uint64_t min_per_thread = 1u << 16;  // below this a thread costs more than it saves
    const uint64_t nt = std::min<uint64_t>(threads, (n + min_per_thread - 1) / min_per_thread);
    if (nt <= 1) return fn(ascii, 0, n, L, n, out);
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    const uint64_t per = (n + nt - 1) / nt;
    uint64_t started = 1;  // ranges [1, started) run on their own threads
    try {
        for (; started < nt; started++)
            pool.emplace_back(fn, ascii, std::min(n, started * per), std::min(n, (started + 1) * per), L, n, out);
    } catch (...) {  // thread limit reached: this thread takes what is left
    }
    fn(ascii, 0, std::min(n, per), L, n, out);
    for (uint64_t t = started; t < nt; t++) fn(ascii, std::min(n, t * per), std::min(n, (t + 1) * per), L, n, out);
    for (auto &th : pool) th.join();
}

void pack_range(const uint8_t *ascii, uint64_t begin, uint64_t end, uint32_t L, uint64_t total, uint64_t *out) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2") && !no_avx2_hook();
    if (have_avx2) return pack_range_avx2(ascii, begin, end, L, total, out);
#endif
    (void)total;
    pack_range_scalar(ascii, begin, end, L, out);
}

#if defined(__x86_64__)
// Dense records straight from the vector masks of pack_range_avx2: a plain read is b1 | b2 << L, stored with one
// overlapping 8-byte store; anything else is an exception whose compact record the scalar packer forms.  Handles the
// reads whose 32-byte load stays inside the buffer and whose 8-byte store stays inside the range ([begin, fast_end)).
__attribute__((target("avx2"))) void dense_range_avx2(const uint8_t *ascii, uint64_t begin, uint64_t fast_end, uint32_t L, uint8_t *out,
                                                      uint32_t *idx, uint64_t *rec_out, uint64_t cap, uint64_t *count) {
    const uint32_t B = SGDX_DENSE_BYTES(L);
    const __m256i lut = _mm256_setr_epi8(1, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1,
                                         1, 'A', 1, 'C', 'T', 1, 1, 'G', 1, 1, 1, 1, 1, 1, 1, 1);
    const __m256i nib = _mm256_set1_epi8(0x0F);
    const uint32_t m = (1u << L) - 1u;
    uint64_t k = *count;
    const uint8_t *src = ascii + begin * L;
    uint8_t *dst = out + begin * B;
    for (uint64_t r = begin; r < fast_end; r++, src += L, dst += B) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src));
        const uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 6));
        const uint32_t b2 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(x, 5));
        const uint32_t ok = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, _mm256_and_si256(x, nib)), x));
        uint64_t d = (uint64_t)(b1 & m) | ((uint64_t)(b2 & m) << L);
        if (__builtin_expect((ok & m) != m, 0)) {
            if (k < cap) {  // the compact record, as pack_range_avx2 forms it
                const uint32_t isn = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('N')));
                const uint32_t inv = ~ok & m;
                idx[k] = (uint32_t)r;
                rec_out[k] = (uint64_t)((b1 & ok & m) | (inv & ~isn)) | ((uint64_t)(b2 & ok & m) << kField) | ((uint64_t)inv << (2 * kField));
            }
            k++;
            d = 0;  // placeholder of an exception
        }
        memcpy(dst, &d, 8);
    }
    *count = k;
}
#endif

// Dense records (include/sgdx.h): packed records of a tile (hot in L1) -> 2 L bits per read + the exception list.
// Exceptions go to idx/rec[*count ...] while there is room (cap); *count keeps counting beyond it.  Records of the
// list: compact 64-bit records up to 21 bases, 16-byte sgdx_read records above.
void dense_range(const uint8_t *ascii, uint64_t begin, uint64_t end, uint32_t L, uint64_t total, uint8_t *out, uint32_t *idx,
                 void *rec_any, uint64_t cap, uint64_t *count) {
    const uint32_t B = SGDX_DENSE_BYTES(L);
    if (L > kField) {  // 22..32 bases: from the 16-byte records
        sgdx_read *rec_out = static_cast<sgdx_read *>(rec_any);
        sgdx_read tile[256];
        uint64_t k = *count;
        for (uint64_t r0 = begin; r0 < end; r0 += 256) {
            const uint64_t cnt = std::min<uint64_t>(256, end - r0);
            planes_range(ascii + r0 * L, 0, cnt, L, total - r0, tile);
            const uint64_t fast = std::min<uint64_t>(cnt, end > r0 + 8 ? end - 8 - r0 : 0);  // reads with r + 8 < end
            uint8_t *dst = out + r0 * B;
            for (uint64_t i = 0; i < cnt; i++, dst += B) {
                const sgdx_read v = tile[i];
                uint64_t d = (uint64_t)v.lo | ((uint64_t)v.hi << L);
                if (__builtin_expect((v.nmask | v.xmask) != 0u, 0)) {
                    if (k < cap) {
                        idx[k] = (uint32_t)(r0 + i);
                        rec_out[k] = v;
                    }
                    k++;
                    d = 0;
                }
                if (i < fast || B == 8) memcpy(dst, &d, 8);
                else memcpy(dst, &d, B);
            }
        }
        *count = k;
        return;
    }
    uint64_t *rec_out = static_cast<uint64_t *>(rec_any);
    const uint64_t fmask = (1ull << L) - 1ull;
    uint64_t tile[512];
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2") && !no_avx2_hook();
    if (have_avx2) {  // the reads whose 32-byte load and 8-byte store are safe, in one fused pass; the rest below
        uint64_t fast_end = end > begin + 8 ? end - 8 : begin;
        const uint64_t bytes = total * L;
        while (fast_end > begin && (fast_end - 1) * L + 32 > bytes) fast_end--;
        dense_range_avx2(ascii, begin, fast_end, L, out, idx, rec_out, cap, count);
        begin = fast_end;
    }
#endif
    uint64_t k = *count;
    for (uint64_t r0 = begin; r0 < end; r0 += 512) {
        const uint64_t cnt = std::min<uint64_t>(512, end - r0);
        pack_range(ascii + r0 * L, 0, cnt, L, total - r0, tile);
        // one 8-byte store per read; what it spills past its B bytes is overwritten by the next reads of this range,
        // and the last 8 reads of a range are copied exactly (the next range may belong to another thread)
        const uint64_t fast = std::min<uint64_t>(cnt, end > r0 + 8 ? end - 8 - r0 : 0);  // reads with r + 8 < end
        uint8_t *dst = out + r0 * B;
        for (uint64_t i = 0; i < cnt; i++, dst += B) {
            const uint64_t rec = tile[i];
            uint64_t d = (rec & fmask) | (((rec >> kField) & fmask) << L);
            if (__builtin_expect((rec >> (2u * kField)) != 0, 0)) {
                if (k < cap) {
                    idx[k] = (uint32_t)(r0 + i);
                    rec_out[k] = rec;
                }
                k++;
                d = 0;  // placeholder of an exception
            }
            if (i < fast) memcpy(dst, &d, 8);
            else memcpy(dst, &d, B);
        }
    }
    *count = k;
}

}  // namespace

extern "C" int sgdx_pack_dense_host(const uint8_t *ascii, uint64_t n, uint32_t barcode_len, uint8_t *out_dense,
                                    uint32_t *exc_index, void *exc_record, uint64_t exc_cap, uint64_t *n_exc,
                                    uint32_t threads) {
    return sgdx::guarded("sgdx_pack_dense_host", [&]() -> int {
    if (n_exc) *n_exc = 0;
    if (barcode_len < 1 || barcode_len > SGDX_MAX_PACKED_LEN)
        return sgdx::fail(SGDX_EINVAL, "dense records hold 1..%u bases, barcode_len is %u", SGDX_MAX_PACKED_LEN, barcode_len);
    const size_t esz = SGDX_DENSE_EXC_BYTES(barcode_len);
    if (!n_exc || (n && (!ascii || !out_dense))) return sgdx::fail(SGDX_EINVAL, "sgdx_pack_dense_host: null buffer");
    if (n >> 32) return sgdx::fail(SGDX_EINVAL, "sgdx_pack_dense_host: %llu reads, at most 2^32 - 1 per batch", (unsigned long long)n);
    if (exc_cap && (!exc_index || !exc_record)) return sgdx::fail(SGDX_EINVAL, "sgdx_pack_dense_host: null exception buffer");
    if (threads == 0) threads = 1;
    const uint64_t min_per_thread = 1u << 16;
    const uint64_t nt = std::max<uint64_t>(1, std::min<uint64_t>(threads, (n + min_per_thread - 1) / min_per_thread));
    uint64_t total = 0;
    if (nt == 1) {  // (the batching stage's 1000-read pushes: no allocation, exceptions straight into the caller's arrays)
        dense_range(ascii, 0, n, barcode_len, n, out_dense, exc_index, exc_record, exc_cap, &total);
    } else {
        struct Part {
            std::vector<uint32_t> idx;
            std::vector<uint8_t> rec;  // esz bytes per exception
            uint64_t count = 0;
        };
        std::vector<Part> parts(nt);
        const uint64_t per = (n + nt - 1) / nt;
        auto work = [&](uint64_t t) {
            const uint64_t b0 = std::min(n, t * per), b1 = std::min(n, (t + 1) * per);
            Part &p = parts[t];
            uint64_t cap = std::max<uint64_t>(1024, (b1 - b0) / 16);
            for (;;) {  // exceptions are few; if a range holds more than its list takes, once more with room for all
                p.idx.resize(cap);
                p.rec.resize(cap * esz);
                p.count = 0;
                dense_range(ascii, b0, b1, barcode_len, n, out_dense, p.idx.data(), p.rec.data(), cap, &p.count);
                if (p.count <= cap) break;
                cap = p.count;
            }
        };
        std::vector<std::thread> pool;
        uint64_t started = 1;
        try {
            for (; started < nt; started++) pool.emplace_back(work, started);
        } catch (...) {  // thread limit reached: this thread takes what is left
        }
        work(0);
        for (uint64_t t = started; t < nt; t++) work(t);
        for (auto &th : pool) th.join();
        for (const auto &p : parts) total += p.count;
        if (total <= exc_cap) {
            uint64_t at = 0;
            for (const auto &p : parts) {  // thread order = ascending read order
                if (p.count) {
                    memcpy(exc_index + at, p.idx.data(), p.count * sizeof(uint32_t));
                    memcpy(static_cast<uint8_t *>(exc_record) + at * esz, p.rec.data(), p.count * esz);
                }
                at += p.count;
            }
        }
    }
    *n_exc = total;
    if (total > exc_cap)
        return sgdx::fail(SGDX_ENOMEM, "sgdx_pack_dense_host: %llu exceptions, room for %llu", (unsigned long long)total, (unsigned long long)exc_cap);
    return SGDX_OK;
    });
}

extern "C" int sgdx_pack_compact_host(const uint8_t *ascii, uint64_t n, uint32_t barcode_len, uint64_t *out,
                                      uint32_t threads) {
    return sgdx::guarded("sgdx_pack_compact_host", [&]() -> int {
    if (barcode_len < 1 || barcode_len > SGDX_COMPACT_MAX_LEN)
        return sgdx::fail(SGDX_EINVAL, "compact records hold 1..%u bases, barcode_len is %u", SGDX_COMPACT_MAX_LEN, barcode_len);
    if (n && (!ascii || !out)) return sgdx::fail(SGDX_EINVAL, "sgdx_pack_compact_host: null buffer");
    run_threads(pack_range, ascii, n, barcode_len, out, threads);
    return SGDX_OK;
    });
}

extern "C" int sgdx_pack_host(const uint8_t *ascii, uint64_t n, uint32_t barcode_len, sgdx_read *out, uint32_t threads) {
    return sgdx::guarded("sgdx_pack_host", [&]() -> int {
    if (barcode_len < 1 || barcode_len > SGDX_MAX_PACKED_LEN)
        return sgdx::fail(SGDX_EINVAL, "barcode_len %u outside 1..%u", barcode_len, SGDX_MAX_PACKED_LEN);
    if (n && (!ascii || !out)) return sgdx::fail(SGDX_EINVAL, "sgdx_pack_host: null buffer");
    run_threads(planes_range, ascii, n, barcode_len, out, threads);
    return SGDX_OK;
    });
}
