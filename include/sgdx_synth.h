/*
 * sgdx_synth.h -- seeded synthetic input generator, measurement/test support.
 * The same counter-based generator runs on the host (C++) and on the device (CUDA) and produces
 * byte-identical reads for any (seed, read id), so any shard of a workload can be regenerated on
 * the CPU for the oracle.  Not part of the reference's interface.
 */
#ifndef SGDX_SYNTH_H
#define SGDX_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgdx_synth_params {
    uint64_t seed;
    uint32_t num_samples;  /* S */
    uint32_t barcode_len;  /* L */
    uint32_t index1_len;   /* 0 = single index (no hop reads are generated) */
    uint32_t zipf;         /* 1 = skewed sample weights (k ~ S*u^2), 0 = uniform */
    /* probabilities as fractions of 2^32 */
    uint32_t p_true;       /* read drawn from one sample (then per-base substitutions) */
    uint32_t p_hop;        /* index1 of sample a + index2 of sample b != a */
    uint32_t p_sub;        /* per-base substitution, applied to true/hop reads */
    uint32_t p_n;          /* per-base no-call injection, all reads */
    uint32_t p_x;          /* per-base foreign byte ('R') injection, all reads */
} sgdx_synth_params;

/* Expected-barcode table: S distinct barcodes; for dual index each half set has pairwise Hamming
 * distance >= min_dist (greedy accept), single index likewise on the whole barcode. out = S*L bytes. */
int sgdx_synth_table(uint64_t seed, uint32_t num_samples, uint32_t barcode_len, uint32_t index1_len,
                     uint32_t min_dist, uint8_t *out);
/* reads [first, first+n) as n*L ASCII bytes */
int sgdx_synth_reads_host(const sgdx_synth_params *p, const uint8_t *table, uint64_t first, uint64_t n,
                          uint8_t *out, int threads);
int sgdx_synth_reads_device(const sgdx_synth_params *p, const uint8_t *d_table, uint64_t first, uint64_t n,
                            uint8_t *d_out, void *cuda_stream);


/* Test support for the perfect-hash kernel (SGDX_PATH_PERFECT_HASH): builds that kernel's tables for `cfg` on the
 * host -- the same code sgdx_create runs -- and answers n compact records with the kernel's fast-path arithmetic,
 * so CPU tests can hold the tables against the oracle.  Records with invalid positions (general path on the
 * device) come back as idx 0xFFFF / dist 0xFE.  info[9] (nullable) = {tables exist, number of keys, bucket bits,
 * slot bits, single-invalid-base shortcut valid, min pairwise table distance, radius of the first level, radius of the
 * second level (0 = none), number of second-level keys}.  It is not an assignment path: it
 * ignores no-calls, the hop checker and the counters. */
struct sgdx_config;
struct sgdx_read;
int sgdx_ph_probe_host(const struct sgdx_config *cfg, const uint8_t *barcodes_ascii, const uint64_t *records,
                       uint64_t n, uint16_t *out_idx, uint8_t *out_dist, uint32_t *info);
/* The same for 16-byte sgdx_read records (barcodes of up to 32 bases).  In both, a read that neither level decides
 * (their radius is smaller than max_mismatches: the kernel's general path searches the seed index) comes back as
 * idx 0xFFFF / dist 0xFD. */
int sgdx_ph_probe_host_wide(const struct sgdx_config *cfg, const uint8_t *barcodes_ascii, const struct sgdx_read *records,
                            uint64_t n, uint16_t *out_idx, uint8_t *out_dist, uint32_t *info);

#ifdef __cplusplus
}
#endif
#endif
