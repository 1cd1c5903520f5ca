// Host packers (csrc/sgdx_hostpack.cpp) under AddressSanitizer + UBSan: exact-size heap buffers for every barcode
// length and a range of batch sizes, so any byte the 32-byte vector loads read past the end of the input is flagged;
// also checks that the 8-byte, the 16-byte and the dense record of every read agree.  Built and run by
// tests/test_compact_cpu.py::test_host_packers_are_clean_under_asan (no GPU, no libsgdx.so: the source is compiled in).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/sgdx.h"
namespace sgdx { int fail(int code, const char *, ...) { return code; } }
int main() {
    const char letters[] = "ACGTNacx";
    for (uint32_t L = 1; L <= 32; L++)
        for (uint64_t n : {0ull, 1ull, 2ull, 3ull, 7ull, 64ull, 1000ull, 70000ull, 200001ull}) {
            // exact-size heap buffers: ASAN flags any byte read past the end
            uint8_t *in = (uint8_t *)malloc(n * L ? n * L : 1);
            for (uint64_t i = 0; i < n * L; i++) in[i] = letters[(i * 2654435761u >> 7) % 8];
            sgdx_read *o16 = (sgdx_read *)malloc(n ? n * sizeof(sgdx_read) : 1);
            if (sgdx_pack_host(in, n, L, o16, 3)) return 1;
            if (L <= 21) {
                uint64_t *o8 = (uint64_t *)malloc(n ? n * 8 : 1);
                if (sgdx_pack_compact_host(in, n, L, o8, 3)) return 2;
                for (uint64_t r = 0; r < n; r++) {  // the two formats agree
                    uint64_t rec = o8[r];
                    uint32_t lo = rec & 0x1FFFFF, hi = (rec >> 21) & 0x1FFFFF, inv = (rec >> 42) & 0x1FFFFF;
                    if ((lo & ~inv) != o16[r].lo || hi != o16[r].hi || (inv & ~lo) != o16[r].nmask || (inv & lo) != o16[r].xmask) return 3;
                }
                // dense records: exact-size output (n * B bytes), exceptions agree with the compact records
                const uint32_t B = SGDX_DENSE_BYTES(L);
                uint8_t *od = (uint8_t *)malloc(n * B ? n * B : 1);
                uint32_t *ei = (uint32_t *)malloc(n ? n * 4 : 1);
                uint64_t *er = (uint64_t *)malloc(n ? n * 8 : 1);
                uint64_t ne = 0;
                if (sgdx_pack_dense_host(in, n, L, od, ei, er, n, &ne, 3)) return 4;
                uint64_t e = 0;
                for (uint64_t r = 0; r < n; r++) {
                    const uint64_t rec = o8[r];
                    if (rec >> 42) {
                        if (e >= ne || ei[e] != r || er[e] != rec) return 5;
                        e++;
                    } else {
                        uint64_t d = 0;
                        memcpy(&d, od + r * B, B);
                        if (d != ((rec & 0x1FFFFF) | (((rec >> 21) & 0x1FFFFF) << L))) return 6;
                    }
                }
                if (e != ne) return 7;
                free(od);
                free(ei);
                free(er);
                free(o8);
            } else {  // 22..32 bases: dense records against the 16-byte records, exceptions are 16-byte records
                const uint32_t B = SGDX_DENSE_BYTES(L);
                uint8_t *od = (uint8_t *)malloc(n * B ? n * B : 1);
                uint32_t *ei = (uint32_t *)malloc(n ? n * 4 : 1);
                sgdx_read *er = (sgdx_read *)malloc(n ? n * sizeof(sgdx_read) : 1);
                uint64_t ne = 0, e = 0;
                if (sgdx_pack_dense_host(in, n, L, od, ei, er, n, &ne, 3)) return 8;
                for (uint64_t r = 0; r < n; r++) {
                    if (o16[r].nmask | o16[r].xmask) {
                        if (e >= ne || ei[e] != r || memcmp(&er[e], &o16[r], sizeof(sgdx_read))) return 9;
                        e++;
                    } else {
                        uint64_t d = 0;
                        memcpy(&d, od + r * B, B);
                        if (d != ((uint64_t)o16[r].lo | ((uint64_t)o16[r].hi << L))) return 10;
                    }
                }
                if (e != ne) return 11;
                free(od);
                free(ei);
                free(er);
            }
            free(o16);
            free(in);
        }
    puts("ok");
    return 0;
}
