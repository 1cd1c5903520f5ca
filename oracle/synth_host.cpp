// synth_host.cpp -- the seeded synthetic generator for the CPU legs of bench.py.
// Test / measurement infrastructure, like the rest of oracle/: the generator's source is
// singular-demux_b200/csrc/sgdx_synth_core.h (shared with the product library so that both sides draw the same
// reads), compiled here into liboracle_synth.so so that `bench.py --impl reference` never loads libsgdx.so.
#include "../singular-demux_b200/csrc/sgdx_synth_core.h"

extern "C" int orc_synth_table(uint64_t seed, uint32_t S, uint32_t L, uint32_t i1, uint32_t min_dist, uint8_t *out) {
    return sgdx::synth_table_impl(seed, S, L, i1, min_dist, out);
}

extern "C" int orc_synth_reads_host(const sgdx_synth_params *p, const uint8_t *table, uint64_t first, uint64_t n, uint8_t *out,
                                    int threads) {
    return sgdx::synth_reads_host_impl(p, table, first, n, out, threads);
}
