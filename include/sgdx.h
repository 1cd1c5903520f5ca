/*
 * sgdx.h -- C ABI of the B200-native sample-barcode assignment path for sgdemux.
 *
 * This is the drop-in boundary.  The reference (Rust, /root/reference/src/lib) has no FFI; the seams
 * this ABI replaces are cited per entry point:
 *
 *   per-read seam   trait Matcher { fn find(&self, barcode: Vec<u8>) -> MatchResult }   matcher.rs:75-77
 *                   CachedHammingDistanceMatcher::new / PreComputeMatcher::new           matcher.rs:87-96,143-151
 *   call site       Demultiplexer::demultiplex, N gate + find + index + counters + hop   demux.rs:524-556,581
 *   selection       PreCompute for <= 12 bp, cached Hamming otherwise                    run.rs:196-236
 *   counters        DemuxedGroupSampleMetrics::update_with_match                         metrics.rs:428-437
 *                   SampleBarcodeHopChecker::try_new / Tracker::check_barcode            metrics.rs:575-614,658-676
 *                   DemuxedGroupMetrics::update_with (merge by sum)                      metrics.rs:315-328
 *
 * Conventions: every function returns 0 on success and a negative SGDX_E* code on failure (the text
 * is available from sgdx_last_error(), thread-local); no exception or abort crosses the boundary; the
 * caller owns every host buffer, the library owns device memory.  A handle may be used from several
 * host threads at once as long as each thread uses its own `slot` (mirrors `M: Matcher + Send + Sync`,
 * demux.rs:422).  There is no CPU fallback: without a CUDA device sgdx_create fails.
 *
 * Barcodes are passed the way the reference passes them: raw ASCII bytes, all sample-barcode (B)
 * segments of a template concatenated in FASTQ order (demux.rs:404-405,504-522), exactly
 * barcode_len bytes per read (run.rs:109-119 guarantees it outside --sample-barcode-in-fastq-header).
 * Barcodes of at most 21 bases may instead be passed as one 64-bit compact record each (2-bit base planes +
 * invalid plane, "Compact records" below): same results, 8 instead of barcode_len bytes per read.
 */
#ifndef SGDX_H
#define SGDX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGDX_OK 0
#define SGDX_EINVAL (-1)   /* bad argument / unsupported configuration */
#define SGDX_ECUDA (-2)    /* CUDA runtime error (no device, launch failure, ...) */
#define SGDX_ENOMEM (-3)
#define SGDX_ESTATE (-4)   /* call made in the wrong state (e.g. unmatched table enabled twice, batcher results not taken) */

/* matcher kinds: MatcherKind, matcher.rs:49-53 */
#define SGDX_MATCHER_AUTO 0u              /* run.rs:196 rule: barcode_len <= 12 -> PreCompute, else CachedHamming */
#define SGDX_MATCHER_CACHED_HAMMING 1u    /* rule set A.2 (matcher.rs:273-348) */
#define SGDX_MATCHER_PRECOMPUTE 2u        /* rule set A.3 (matcher.rs:121-262) */

/* kernel variants (both are hand-written sm_100a kernels with identical results) */
#define SGDX_KERNEL_AUTO 0u
#define SGDX_KERNEL_DIRECT 1u     /* per (read, sample) XOR/OR/POPC scan with best/second-best keys */
#define SGDX_KERNEL_SEEDED 2u     /* pigeonhole seed index -> verify only the candidate samples; falls back to DIRECT when the seed index does not apply */

#define SGDX_MAX_BARCODE_LEN 64u  /* ASCII entry points; 33..64 bases run sgdx_match_long (each index half <= 32 bases) */
#define SGDX_MAX_PACKED_LEN 32u   /* sgdx_read records, the unmatched table and the synthetic generator */
#define SGDX_MAX_SAMPLES 8192u
#define SGDX_NO_DIST 0xFFu        /* out_dist value of a NoMatch read */

typedef struct sgdx_handle sgdx_handle;

typedef struct sgdx_config {
    uint32_t num_samples;     /* S, Undetermined excluded (run.rs:201,220); 1..SGDX_MAX_SAMPLES */
    uint32_t barcode_len;     /* L, 1..SGDX_MAX_BARCODE_LEN (records and the unmatched table: 1..SGDX_MAX_PACKED_LEN) */
    uint32_t max_mismatches;  /* --allowed-mismatches, opts.rs:118 */
    uint32_t min_delta;       /* --min-delta, opts.rs:123 */
    uint32_t free_ns;         /* --free-ns, opts.rs:127 (ignored by PreCompute, as in the reference) */
    int32_t  max_no_calls;    /* --max-no-calls, opts.rs:135; -1 = None */
    uint32_t matcher;         /* SGDX_MATCHER_* */
    uint32_t index1_len;      /* hop checker (metrics.rs:587-595): summed B lengths of the 1st barcode read */
    uint32_t index2_len;      /*   ... of the 2nd; both 0 = fewer/more than two barcode reads -> checker off */
    int32_t  device;          /* CUDA device ordinal */
    uint32_t slots;           /* independent in-flight batches (one stream each); 0 -> 2 */
    uint32_t kernel;          /* SGDX_KERNEL_* */
} sgdx_config;

/* 16-byte packed read record (2-bit base planes + no-call mask + foreign-byte mask); bit j = base j */
typedef struct sgdx_read {
    uint32_t lo, hi;  /* code = (ascii >> 1) & 3 : A=0 C=1 T=2 G=3 ; lo = bit 0, hi = bit 1 */
    uint32_t nmask;   /* byte == 'N' */
    uint32_t xmask;   /* byte not in {A,C,G,T,N}: mismatches every sample, is not a no-call */
} sgdx_read;

/* cumulative counters of a handle; per_sample is [(S+1)][3] = {total, perfect, one_mismatch},
 * Undetermined last (metrics.rs:407-437; demux.rs:259) */
typedef struct sgdx_totals {
    uint64_t total_templates;   /* demux.rs:581 */
    uint64_t matched;           /* reads assigned to a real sample */
    uint64_t undetermined;
    uint64_t hop_reads;         /* NoMatch reads counted by the hop tracker */
} sgdx_totals;

const char *sgdx_last_error(void);
const char *sgdx_version(void);

/* Matcher construction: replaces CachedHammingDistanceMatcher::new / PreComputeMatcher::new plus
 * SampleBarcodeHopChecker::try_new.  barcodes_ascii = S*L bytes, upper-case ACGT only
 * (sample_metadata.rs:219-226); anything else is SGDX_EINVAL. */
int sgdx_create(const sgdx_config *cfg, const uint8_t *barcodes_ascii, sgdx_handle **out);
void sgdx_destroy(sgdx_handle *h);
int sgdx_get_config(const sgdx_handle *h, sgdx_config *effective); /* AUTO fields resolved */

/* pinned host memory for batch buffers (plain memory also works, through an internal staging copy) */
int sgdx_alloc_host(size_t bytes, void **out);
int sgdx_free_host(void *p);

/* Batched Matcher::find + demux.rs:524-556 counting, HOST buffers.  Asynchronous: queues
 * H2D(n*L bytes) -> assign kernel -> D2H(idx, dist) on the slot's stream and returns; the outputs are
 * valid after sgdx_sync(h, slot).  out_idx[r] in [0,S], S = Undetermined; out_dist[r] = reported
 * hamming_dist (free-N discounted, matcher.rs:301) or SGDX_NO_DIST.  out_dist may be NULL. */
int sgdx_match_batch(sgdx_handle *h, uint32_t slot, const uint8_t *barcodes_ascii, uint64_t n,
                     uint16_t *out_idx, uint8_t *out_dist);
int sgdx_sync(sgdx_handle *h, uint32_t slot, float *device_ms /* nullable: kernel time of last batch */);

/* Same, buffers already resident on the handle's device (no copies).  When barcode_len is not a multiple of 4
 * or the pointer is not 4-byte aligned the kernels load the aligned 32-bit words that contain a read, so up to
 * 3 bytes before the first and after the last read of the buffer are read (never used): keep the buffer inside
 * one allocation. */
int sgdx_match_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_barcodes_ascii, uint64_t n,
                      uint16_t *d_out_idx, uint8_t *d_out_dist);
/* Two-stage form: pack kernel (ASCII -> 16-byte records, 128-bit stores) and match on packed records. */
int sgdx_pack_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_barcodes_ascii, uint64_t n,
                     sgdx_read *d_packed);
int sgdx_match_packed_device(sgdx_handle *h, uint32_t slot, const sgdx_read *d_packed, uint64_t n,
                             uint16_t *d_out_idx, uint8_t *d_out_dist);
/* The packed form from HOST buffers (SURVEY Appendix B's `sgdx_pack` + packed `sgdx_match_batch`): sgdx_pack_host
 * converts n ASCII barcodes of any barcode_len <= 32 into 16-byte sgdx_read records (host code, `threads` workers,
 * AVX2 when the CPU has it; same records as sgdx_pack_device), sgdx_match_packed_batch is sgdx_match_batch for such
 * a buffer (16 instead of barcode_len bytes per read over PCIe: worth it above 16 bases; for <= 21 bases the compact
 * records below are smaller).  The unmatched table (sgdx_unmatched_enable) is keyed from the records; a NoMatch barcode
 * with a byte that is none of A/C/G/T/N cannot be keyed from a record (the byte itself is not kept) and is counted
 * in sgdx_unmatched_info.dropped_reads -- the ASCII entry points count those as well. */
int sgdx_pack_host(const uint8_t *barcodes_ascii, uint64_t n, uint32_t barcode_len, sgdx_read *out_records,
                   uint32_t threads /* 0 -> 1 */);
int sgdx_match_packed_batch(sgdx_handle *h, uint32_t slot, const sgdx_read *records, uint64_t n,
                            uint16_t *out_idx, uint8_t *out_dist);
/* Compact records: one 64-bit word per read, for barcode_len <= SGDX_COMPACT_MAX_LEN -- BASELINE.json's
 * "2-bit base planes plus an N mask", SURVEY 8(d)'s algorithmic input of ceil(3L/8) bytes per read:
 *   bits  0..20  lo plane: bit j = bit 0 of base j's code, code = (ascii >> 1) & 3 (A0 C1 T2 G3)
 *   bits 21..41  hi plane: bit j = bit 1 of the code
 *   bits 42..62  invalid plane: base j is not A/C/G/T; there lo = 0 means 'N' (a no-call, demux.rs:78-80),
 *                lo = 1 any other byte (mismatches every sample, as in matcher.rs:335-348); hi = 0
 *   bit  63      zero; planes are zero at positions >= barcode_len
 * The batching stage (demux.rs:522 -> :528) writes this instead of ASCII into its pinned buffer:
 * sgdx_pack_compact_host is that pass (host code, `threads` workers, AVX2 when the CPU has it), and
 * sgdx_match_compact_batch is sgdx_match_batch for such a buffer: 8 instead of barcode_len bytes per read
 * cross PCIe, and the kernel (sgdx_match_ph, see SGDX_PATH_PERFECT_HASH) reads four of them with one 256-bit load.
 * Results are identical to the ASCII entry points for every configuration; tables too large for the perfect hash
 * are served through the 16-byte packed form of the general kernels.  Bits above barcode_len in any plane, and
 * bit 63, are ignored.  The unmatched table is keyed from the records (see above for bytes outside A/C/G/T/N). */
#define SGDX_COMPACT_MAX_LEN 21u
/* Dense records: the packed records as they cross PCIe in batches from the host -- 2 bits per base and nothing else,
 * SGDX_DENSE_BYTES(barcode_len) bytes per read (5 bytes at 20 bases instead of 8, 6 at 24 instead of 16), for
 * barcode_len <= SGDX_MAX_PACKED_LEN: bits 0..L-1 the lo plane, bits L..2L-1 the hi plane, little endian, read r at
 * byte r * SGDX_DENSE_BYTES(L).  The invalid plane travels as a list of exceptions: a read that is not made of A/C/G/T
 * only (a no-call, any other byte: about 2 % of real reads) has an arbitrary placeholder in the dense stream and an
 * entry (read index, record) in the list -- the record is a 64-bit compact record for barcode_len <=
 * SGDX_COMPACT_MAX_LEN and a 16-byte sgdx_read above (SGDX_DENSE_EXC_BYTES).  On the device a streaming kernel rebuilds
 * the packed records (sgdx_unpack_dense + sgdx_patch_exceptions) and the assignment runs on those: same results as
 * every other entry point.
 * sgdx_pack_dense_host: out_dense = n * SGDX_DENSE_BYTES(L) bytes; exceptions in ascending read order; if there are
 * more than exc_cap of them the call fails with SGDX_ENOMEM and *n_exc says how many there are.
 * sgdx_match_dense_batch: sgdx_match_batch for such a batch (n < 2^32). */
#define SGDX_DENSE_BYTES(barcode_len) ((2u * (barcode_len) + 7u) / 8u)
#define SGDX_DENSE_EXC_BYTES(barcode_len) ((barcode_len) <= SGDX_COMPACT_MAX_LEN ? 8u : 16u)
int sgdx_pack_dense_host(const uint8_t *barcodes_ascii, uint64_t n, uint32_t barcode_len, uint8_t *out_dense,
                         uint32_t *exc_index, void *exc_record, uint64_t exc_cap, uint64_t *n_exc, uint32_t threads);
int sgdx_match_dense_batch(sgdx_handle *h, uint32_t slot, const uint8_t *dense, uint64_t n, const uint32_t *exc_index,
                           const void *exc_record, uint64_t n_exc, uint16_t *out_idx, uint8_t *out_dist);
int sgdx_pack_compact_host(const uint8_t *barcodes_ascii, uint64_t n, uint32_t barcode_len, uint64_t *out_records,
                           uint32_t threads /* 0 -> 1 */);
int sgdx_pack_compact_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_barcodes_ascii, uint64_t n,
                             uint64_t *d_records);
int sgdx_match_compact_batch(sgdx_handle *h, uint32_t slot, const uint64_t *records, uint64_t n,
                             uint16_t *out_idx, uint8_t *out_dist);
int sgdx_match_compact_device(sgdx_handle *h, uint32_t slot, const uint64_t *d_records, uint64_t n,
                              uint16_t *d_out_idx, uint8_t *d_out_dist);
/* Run the slot on a caller-owned stream (cudaStream_t passed as void*); NULL restores the slot's own. */
int sgdx_set_stream(sgdx_handle *h, uint32_t slot, void *cuda_stream);
/* Name of the assignment kernel this handle launches for ASCII input ("sgdx_match_direct",
 * "sgdx_match_seeded" or "sgdx_match_exact_first"); static storage. */
const char *sgdx_kernel_name(const sgdx_handle *h);
/* ... and for compact records ("sgdx_match_ph"; "sgdx_match_phx", its general form, when the key set stops short of
 * max_mismatches; "sgdx_match_single" for one sample without hop checker; or the ASCII name when the records are expanded for the general
 * kernels; "" when barcode_len > SGDX_COMPACT_MAX_LEN). */
const char *sgdx_compact_kernel_name(const sgdx_handle *h);
/* ... and for 16-byte sgdx_read records ("sgdx_match_phx" for barcodes of 22..32 bases whose tables allow it, else
 * "sgdx_match_seeded" / "sgdx_match_direct"). */
const char *sgdx_packed_kernel_name(const sgdx_handle *h);
/* Which accelerations this handle's tables allow (they never change results, only speed). */
#define SGDX_PATH_SEED_INDEX 1u   /* pigeonhole seed index exists (T+1 <= L, fits shared memory) */
#define SGDX_PATH_EXACT 2u        /* exact-match front end (min pairwise table distance > min_delta) */
#define SGDX_PATH_NEIGHBOURS 4u   /* one-substitution neighbour table (min pairwise distance > min_delta + 2) */
#define SGDX_PATH_HOP_DIRECT 8u   /* direct-address hop tables (both index halves <= 11 bases) */
/* 16u, 32u: flags of the round-1 compact kernel, retired (never set) */
#define SGDX_PATH_PERFECT_HASH 64u /* packed records run sgdx_match_ph (compact records for barcode_len <= 21, 16-byte
                                     sgdx_read records above): the A/C/G/T strings within a radius of the barcodes (up to
                                     max_mismatches, as many as ~28 000 slots hold) sit in a perfect hash with their
                                     precomputed verdicts; SGDX_DISABLE_PH=1 at sgdx_create turns it off */
#define SGDX_PATH_SHORT_SEED 128u  /* the perfect hash stops short of max_mismatches (C3: 1536 x 24 bases at 3 mismatches) and
                                     what it leaves undecided searches a seed index of max_mismatches + 1 chunks whose
                                     verdict is checked against the table's minimum pairwise distance */
uint32_t sgdx_path_flags(const sgdx_handle *h);
/* Number of kernel launches issued through this handle so far. */
uint64_t sgdx_launch_count(const sgdx_handle *h);

/* Counters (device-resident, cumulative over all slots; synchronises the device).
 * per_sample: (S+1)*3 u64, nullable.  Replaces reading DemuxedGroupMetrics after the fold/reduce. */
int sgdx_counters(sgdx_handle *h, uint64_t *per_sample, sgdx_totals *totals);
int sgdx_reset_counters(sgdx_handle *h);
/* Hop tracker contents (metrics.rs:658-676): number of distinct hop barcodes, then export of
 * (concatenated observed barcode [L bytes], count) pairs in unspecified order. */
int sgdx_hop_count(sgdx_handle *h, uint64_t *n_keys);
int sgdx_hop_export(sgdx_handle *h, uint8_t *keys /* cap*L */, uint64_t *counts /* cap */, uint64_t cap,
                    uint64_t *n_written);

/* ---- host batching stage (the new stage between segment extraction demux.rs:522 and assignment
 * demux.rs:528): accumulates the barcodes of many 1000-read chunks into pinned super-batches, keeps
 * `slots` batches in flight and hands results back in submission order. */
typedef struct sgdx_batcher sgdx_batcher;
int sgdx_batcher_create(sgdx_handle *h, uint64_t batch_reads, sgdx_batcher **out);
void sgdx_batcher_destroy(sgdx_batcher *b);
/* append n barcodes (n*L ASCII bytes); launches a batch whenever one fills.  At most 2 x slots batches may wait for
 * sgdx_batcher_next (each holds pinned result buffers): beyond that the push stops with SGDX_ESTATE.  On any failure
 * *consumed (if not NULL) says how many of the n barcodes were taken, so that the caller can resume behind them
 * instead of pushing them twice; sgdx_batcher_push is the same call without that count. */
int sgdx_batcher_push_n(sgdx_batcher *b, const uint8_t *barcodes_ascii, uint64_t n, uint64_t *consumed);
int sgdx_batcher_push(sgdx_batcher *b, const uint8_t *barcodes_ascii, uint64_t n);
/* launch the partially filled batch, if any */
int sgdx_batcher_flush(sgdx_batcher *b);
/* on = 1: sgdx_batcher_push packs each barcode into a compact record as it copies it into the pinned buffer, and
 * batches go up as 8 bytes per read; on = 2: as dense records (2 bits per base + the exception list, see below), about
 * 5 bytes per read at 20 bases (barcode_len <= SGDX_COMPACT_MAX_LEN).  Call before the first push or after the last
 * result was taken. */
int sgdx_batcher_set_compact(sgdx_batcher *b, int on);
/* Oldest unconsumed batch: blocks until it is done; pointers stay valid until the next call of
 * sgdx_batcher_next/destroy.  *n == 0 when nothing is pending. */
int sgdx_batcher_next(sgdx_batcher *b, const uint16_t **idx, const uint8_t **dist, uint64_t *n);

/* ==== the stages either side of the assignment; all device work, no CPU fallback ==== */

/* ---- 1. most-frequent unmatched barcodes on the device.
 * Replaces UnmatchedCounter (metrics.rs:131-190) and the counter thread it is fed through
 * (metrics.rs:79-100, demux.rs:553-555, run.rs:185-190,262-268): after sgdx_unmatched_enable every batch
 * also counts the concatenated barcode of each NoMatch read in a device hash table (3 bits per base, 128-bit
 * compare-and-swap; barcodes with a byte outside ACGTN go to a side list and are counted on the host).
 * Keys are the concatenated observed barcodes; the host writes them in their delimited form
 * (demux.rs:504-522).  Down-sizing follows the reference's heuristic (metrics.rs:151-175) at batch
 * granularity: when the number of distinct keys has reached max_map_size at a sgdx_sync, the table is cut to
 * its downsize_to most frequent keys.  Below max_map_size distinct keys the counts are exact.  Needs ASCII input. */
typedef struct sgdx_unmatched_info {
    uint64_t distinct_keys;     /* keys in the table now (incl. foreign-byte barcodes) */
    uint64_t unmatched_reads;   /* NoMatch reads seen since enable/reset */
    uint64_t dropped_reads;     /* reads not counted because the table (or the foreign-byte list) was full */
    uint64_t downsizes;         /* how many times the table was cut down */
} sgdx_unmatched_info;
int sgdx_unmatched_enable(sgdx_handle *h, uint64_t max_map_size /* 0 -> 5,000,000 (opts.rs:185) */,
                          uint64_t downsize_to /* 0 -> 5,000 (opts.rs:189) */);
int sgdx_unmatched_info_get(sgdx_handle *h, sgdx_unmatched_info *out);
/* every (key, count) pair, unspecified order; keys = cap*L bytes.  For merging shards on the host. */
int sgdx_unmatched_export(sgdx_handle *h, uint8_t *keys, uint64_t *counts, uint64_t cap, uint64_t *n_written);
/* the n most frequent keys, descending count (to_file, metrics.rs:177-190; ties in byte order here, hash
 * order in the reference).  The count threshold is found from a device-side histogram, so only the selected
 * keys cross PCIe. */
int sgdx_unmatched_top(sgdx_handle *h, uint64_t n, uint8_t *keys /* n*L */, uint64_t *counts /* n */,
                       uint64_t *n_written);
int sgdx_unmatched_downsize(sgdx_handle *h);   /* UnmatchedCounter::downsize now (metrics.rs:166-175) */

/* ---- 2. sample-barcode extraction on the device.
 * Replaces the sample-barcode part of Demultiplexer::process_segments / demultiplex (demux.rs:360-417,
 * 504-522): the B segments of each FASTQ's read, in FASTQ order then segment order, concatenated into the
 * L-byte barcode.  Sources are fixed-stride matrices of bases, one per FASTQ (row r = read r). */
#define SGDX_SEG_TEMPLATE 'T'
#define SGDX_SEG_SAMPLE_BARCODE 'B'
#define SGDX_SEG_MOLECULAR_BARCODE 'M'
#define SGDX_SEG_SKIP 'S'
#define SGDX_SEG_TO_END 0xFFFFFFFFu
typedef struct sgdx_segment {      /* read_structure::ReadSegment as used at demux.rs:380-387 */
    uint32_t source;               /* which FASTQ (index into the source arrays); ignored by single-source calls */
    uint32_t offset;               /* first base of the segment in the read */
    uint32_t length;               /* bases, or SGDX_SEG_TO_END for the trailing `+` segment */
    uint32_t kind;                 /* SGDX_SEG_* */
} sgdx_segment;
#define SGDX_MAX_SOURCES 8u
#define SGDX_MAX_SEGMENTS 32u
int sgdx_extract_device(sgdx_handle *h, uint32_t slot, uint32_t num_sources, const uint8_t *const *d_bases,
                        const uint32_t *strides, const sgdx_segment *segments, uint32_t num_segments, uint64_t n,
                        uint8_t *d_barcodes_out /* n*L */);
/* extraction into the slot's scratch + sgdx_match_device in one call */
int sgdx_match_segments_device(sgdx_handle *h, uint32_t slot, uint32_t num_sources, const uint8_t *const *d_bases,
                               const uint32_t *strides, const sgdx_segment *segments, uint32_t num_segments,
                               uint64_t n, uint16_t *d_out_idx, uint8_t *d_out_dist);

/* ---- 3. base-quality counters per assigned sample.
 * Replaces BaseQualCounter::update (metrics.rs:499-518) over every segment of one FASTQ and the per-sample
 * merge at demux.rs:582-583: for read r with sample index d_idx[r], template (T) bases add to bases / q20 /
 * q30, sample-barcode (B) bases add to index_bases_total_seen / index_bases_qual_sum.  q = byte - 33 in u8
 * arithmetic (a byte below 33 wraps, as in a release build of the reference).  d_quals is a fixed-stride
 * matrix (row r = quality string of read r, 16-byte aligned base); d_lens (nullable) gives each read's
 * length <= stride, which a trailing SGDX_SEG_TO_END segment follows.  A read shorter than its fixed
 * segments (the reference fails on it, demux.rs:380-387) adds nothing and is counted in short_reads.
 * Call once per FASTQ of the batch. */
typedef struct sgdx_base_qual {    /* BaseQualCounter, metrics.rs:468-481 */
    uint64_t q20_bases, q30_bases, index_bases_qual_sum, index_bases_total_seen, bases;
} sgdx_base_qual;
int sgdx_qual_counts_device(sgdx_handle *h, uint32_t slot, const uint8_t *d_quals, uint64_t n, uint32_t stride,
                            const uint32_t *d_lens, const sgdx_segment *segments, uint32_t num_segments,
                            const uint16_t *d_idx);
int sgdx_base_qual_counters(sgdx_handle *h, sgdx_base_qual *per_sample /* S+1, Undetermined last */,
                            uint64_t *short_reads /* nullable */);

/* ---- 4. read routing: a stable partition of the read ids by sample index, so that the host writer
 * (run.rs:254-261, pooled_sample_writer.rs:25-40) receives one contiguous, input-ordered run per sample
 * (what DemuxedGroup::per_sample_reads holds, demux.rs:637-645).  d_order[offsets[s] .. offsets[s+1]) are the
 * ids of sample s's reads in input order; offsets has S+2 entries, offsets[S+1] = n.  n < 2^32. */
int sgdx_route_device(sgdx_handle *h, uint32_t slot, const uint16_t *d_idx, uint64_t n, uint32_t *d_order,
                      uint64_t *d_offsets /* S+2 */);

/* ---- 5. one call per batch: the device-resident form of Demultiplex::demultiplex(&PerFastqRecordSet) ->
 * DemuxedGroup (demux.rs:164-173,445-584) after the header filters: sample-barcode extraction from every FASTQ's
 * reads, assignment with its counters (and the unmatched table when enabled), base-quality counters of every
 * FASTQ that brings qualities, and -- when d_order is given -- the routing of the reads to their samples, all
 * queued on the slot's stream.  Header rewriting, quality masking and the writers stay on the host. */
typedef struct sgdx_fastq_view {   /* the records of one FASTQ of the batch as fixed-stride matrices on the device */
    const uint8_t *d_bases;        /* n * stride */
    const uint8_t *d_quals;        /* n * stride, 16-byte aligned; NULL = no quality counters from this FASTQ */
    const uint32_t *d_lens;        /* per-read length <= stride; NULL = every read is `stride` long */
    uint32_t stride;
} sgdx_fastq_view;
typedef struct sgdx_demux_out {
    uint16_t *d_sample_idx;        /* n; S = Undetermined (demux.rs:535-538) */
    uint8_t *d_hamming_dist;       /* n, nullable */
    uint32_t *d_order;             /* n, nullable: per_sample_reads as one stable permutation ... */
    uint64_t *d_offsets;           /* ... and its S+2 run bounds (required with d_order) */
} sgdx_demux_out;
int sgdx_demultiplex_device(sgdx_handle *h, uint32_t slot, const sgdx_fastq_view *fastqs, uint32_t num_fastqs,
                            const sgdx_segment *segments /* .source = index into fastqs */, uint32_t num_segments,
                            uint64_t n, const sgdx_demux_out *out);

/* ---- measurement support */
/* Integer-pipe micro-benchmark on `device`: sustained 32-bit POPC and LOP3 results per second. */
int sgdx_measure_int_peak(int device, double *popc_per_s, double *lop3_per_s);

#ifdef __cplusplus
}
#endif
#endif /* SGDX_H */
