// bench_batcher.cpp -- the batching stage (sgdx_batcher_*, the new stage between demux.rs:522 and :528) fed the way the
// host pipeline feeds it: T producer threads, each with its own handle + batcher, push 1000-read chunks of barcodes
// that are hot in cache (what the segment extractor has just written) -- once with the plain ASCII copy into the
// pinned buffer, once with the copy replaced by the compact-record packing pass, once by the dense-record one
// (sgdx_batcher_set_compact 1 / 2).  Wall
// clock over push + flush + taking every result; H2D, kernel and D2H included.  Plain C ABI, like a Rust host would.
//
//   g++ -O2 -std=c++17 -I include tools/bench_batcher.cpp -o gpurun_out/bench_batcher -L singular-demux_b200 -lsgdx \
//       -Wl,-rpath,$PWD/singular-demux_b200 -lpthread
//   gpurun_out/bench_batcher [reads per thread = 64000000] [threads = 1,4,8,16]
// Workload: BASELINE config 2 (384 x (10+10) bases, defaults).  Prints one JSON line.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "sgdx.h"
#include "sgdx_synth.h"

#define CK(call)                                                                       \
    do {                                                                               \
        if (int rc__ = (call)) {                                                       \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, sgdx_last_error());   \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

int main(int argc, char **argv) {
    const uint64_t per_thread = argc > 1 ? strtoull(argv[1], nullptr, 10) : 64000000ull;
    std::vector<int> counts;
    {
        std::string s = argc > 2 ? argv[2] : "1,4,8,16";
        for (size_t p = 0; p < s.size();) {
            size_t q = s.find(',', p);
            if (q == std::string::npos) q = s.size();
            counts.push_back(atoi(s.substr(p, q - p).c_str()));
            p = q + 1;
        }
    }
    const uint32_t S = 384, L = 20, chunk = 1000, pool_chunks = 256;
    const uint64_t batch_reads = 2000000;
    std::vector<uint8_t> table((size_t)S * L);
    CK(sgdx_synth_table(0xB202, S, L, 10, 3, table.data()));
    sgdx_synth_params sp{};
    sp.seed = 0x5EEF;
    sp.num_samples = S, sp.barcode_len = L, sp.index1_len = 10, sp.zipf = 1;
    sp.p_true = (uint32_t)(0.92 * 4294967296.0), sp.p_hop = (uint32_t)(0.03 * 4294967296.0);
    sp.p_sub = (uint32_t)(0.005 * 4294967296.0), sp.p_n = (uint32_t)(0.001 * 4294967296.0);
    std::vector<uint8_t> pool((size_t)pool_chunks * chunk * L);  // 5 MB: stays in the producers' cache
    CK(sgdx_synth_reads_host(&sp, table.data(), 0, (uint64_t)pool_chunks * chunk, pool.data(), 4));
    sgdx_config cfg{};
    cfg.num_samples = S, cfg.barcode_len = L, cfg.max_mismatches = 1, cfg.min_delta = 2, cfg.free_ns = 1, cfg.max_no_calls = -1;
    cfg.matcher = SGDX_MATCHER_AUTO, cfg.index1_len = 10, cfg.index2_len = 10, cfg.device = 0, cfg.slots = 2, cfg.kernel = 0;

    printf("{\"workload\": \"C2 384x(10+10), %llu reads per producer thread in %u-read chunks (%u distinct chunks, hot in cache), "
           "%llu-read batches, 2 slots per producer\", \"host_threads\": %u, \"rows\": [",
           (unsigned long long)per_thread, chunk, pool_chunks, (unsigned long long)batch_reads, std::thread::hardware_concurrency());
    bool first_row = true;
    for (int T : counts) {
        printf("%s{\"threads\": %d", first_row ? "" : ", ", T);
        first_row = false;
        for (int compact = 0; compact < 3; compact++) {  // 0 ASCII copy, 1 compact records, 2 dense records
            std::vector<sgdx_handle *> hs(T);
            std::vector<sgdx_batcher *> bs(T);
            for (int i = 0; i < T; i++) {
                CK(sgdx_create(&cfg, table.data(), &hs[i]));
                CK(sgdx_batcher_create(hs[i], batch_reads, &bs[i]));
                CK(sgdx_batcher_set_compact(bs[i], compact));
            }
            std::atomic<uint64_t> matched{0};
            double secs = 0;
            for (int pass = 0; pass < 2; pass++) {  // pass 0 warms up (device staging, pinned result buffers)
                const uint64_t n_reads = pass == 0 ? 4 * batch_reads : per_thread;
                matched = 0;
                auto producer = [&](int i) {
                    const uint16_t *idx;
                    const uint8_t *dist;
                    uint64_t k, pushed = 0, due = 0, got = 0, hits = 0;
                    auto take = [&](bool all) {
                        do {
                            CK(sgdx_batcher_next(bs[i], &idx, &dist, &k));
                            got += k;
                            for (uint64_t r = 0; r < k; r += 4096) hits += idx[r] < S;  // touch the results
                        } while (all && k);
                    };
                    for (uint64_t j = (uint64_t)i * 17; pushed < n_reads; j++) {
                        CK(sgdx_batcher_push(bs[i], pool.data() + (j % pool_chunks) * (size_t)chunk * L, chunk));
                        pushed += chunk, due += chunk;
                        if (due >= 2 * batch_reads) {  // results are taken as they become due, like a writer stage would
                            take(false);
                            due -= batch_reads;
                        }
                    }
                    CK(sgdx_batcher_flush(bs[i]));
                    take(true);
                    if (got != n_reads) {
                        fprintf(stderr, "producer %d: %llu results for %llu reads\n", i, (unsigned long long)got, (unsigned long long)n_reads);
                        exit(1);
                    }
                    matched += hits;
                };
                const auto t0 = std::chrono::steady_clock::now();
                std::vector<std::thread> th;
                for (int i = 0; i < T; i++) th.emplace_back(producer, i);
                for (auto &t : th) t.join();
                secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            }
            const double rps = (double)T * (double)per_thread / secs;
            printf(", \"%s\": {\"reads_per_s\": %.4g, \"seconds\": %.3f, \"pcie_up_gbs\": %.2f}",
                   compact == 2 ? "dense" : (compact ? "compact" : "ascii"), rps, secs,
                   rps * (compact == 2 ? 5.25 : (compact ? 8 : L)) / 1e9);
            for (int i = 0; i < T; i++) {
                sgdx_batcher_destroy(bs[i]);
                sgdx_destroy(hs[i]);
            }
        }
        printf("}");
    }
    printf("]}\n");
    return 0;
}
