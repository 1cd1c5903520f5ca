// sgdx_host.hpp -- C++ host-side mirror of the reference's interface for the assignment path, over the C ABI
// (include/sgdx.h).  The reference is Rust; this image has no cargo/rustc, so the compiled-language host side is
// C++17, header only.  Names, argument meaning and error behaviour follow /root/reference/src/lib:
//
//   MatchResult                       matcher.rs:27-47      MatcherKind, select_matcher      matcher.rs:49-72, run.rs:196-198
//   trait Matcher::find               matcher.rs:75-77      CachedHammingDistanceMatcher     matcher.rs:80-109
//   PreComputeMatcher                 matcher.rs:121-262    DemuxedGroupSampleMetrics        metrics.rs:407-437
//   SampleBarcodeHopTracker           metrics.rs:617-677    DemuxedGroupMetrics::update_with metrics.rs:315-328
//   Demultiplexer's assignment step   demux.rs:524-556,581  (BarcodeAssignmentStage: the batched form a GPU needs)
//   ReadStructure / ReadSegment       opts.rs:84-92 (crate read-structure 0.1.0, used at demux.rs:380-387)
//   BaseQualCounter                   metrics.rs:468-518    UnmatchedCounter / BarcodeCount  metrics.rs:131-206
//
// Errors: the reference returns anyhow::Result from everything around the matcher and find() itself cannot fail;
// here every failed C-ABI call throws sgdx::Error carrying sgdx_last_error().  There is no CPU fallback: without
// libsgdx.so / a CUDA device construction throws.
#pragma once
#include <cstdint>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sgdx.h"

namespace sgdx {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what + ": " + sgdx_last_error()), code(c) {}
};
inline void check(int rc, const char *what) {
    if (rc != SGDX_OK) throw Error(rc, what);
}

inline const char *const UNDETERMINED_NAME = "Undetermined";  // matcher.rs:19

struct SampleMetadata {  // the fields of sample_metadata.rs::SampleMetadata this path reads
    std::string sample_id;
    std::string barcode;
    size_t ordinal = 0;
};

struct MatchResult {  // matcher.rs:27-47
    bool matched = false;
    size_t sample_index = 0;
    size_t hamming_dist = 0;
    std::string barcode;
    static MatchResult Match(size_t sample_index, size_t hamming_dist, std::string barcode) {
        return {true, sample_index, hamming_dist, std::move(barcode)};
    }
    static MatchResult NoMatch(std::string barcode) { return {false, 0, 0, std::move(barcode)}; }
    bool is_match() const { return matched; }
    bool is_no_match() const { return !matched; }
};

enum class MatcherKind { CachedHammingDistance, PreCompute };  // matcher.rs:49-53

inline MatcherKind select_matcher(size_t first_barcode_len, std::optional<MatcherKind> override_matcher = {}) {
    // run.rs:196-198
    if (first_barcode_len <= 12 || override_matcher == MatcherKind::PreCompute) return MatcherKind::PreCompute;
    return MatcherKind::CachedHammingDistance;
}

struct Matcher {  // matcher.rs:75-77
    virtual ~Matcher() = default;
    virtual MatchResult find(std::string barcode) = 0;
};

// A matcher bound to one CUDA device (owns an sgdx handle).
class GpuMatcher : public Matcher {
  public:
    GpuMatcher(const std::vector<SampleMetadata> &samples, MatcherKind kind, size_t max_mismatches, size_t min_delta,
               size_t free_ns, std::optional<size_t> max_no_calls = {},
               std::optional<std::pair<size_t, size_t>> index_lengths = {}, int device = 0, unsigned slots = 2)
        : kind_(kind) {
        if (samples.empty()) throw std::invalid_argument("a GPU matcher needs at least one sample");
        L_ = samples[0].barcode.size();
        std::string table;
        for (const auto &s : samples) {
            if (s.barcode.size() != L_) throw std::invalid_argument("all sample barcodes must have the same length");
            table += s.barcode;
        }
        S_ = samples.size();
        auto clamp = [](size_t v) { return (uint32_t)(v > 0xFFFFFFFFull ? 0xFFFFFFFFull : v); };
        sgdx_config cfg{};
        cfg.num_samples = (uint32_t)S_;
        cfg.barcode_len = (uint32_t)L_;
        cfg.max_mismatches = clamp(max_mismatches);
        cfg.min_delta = clamp(min_delta);
        cfg.free_ns = clamp(free_ns);
        cfg.max_no_calls = max_no_calls ? (int32_t)(*max_no_calls > 0x7FFFFFFF ? 0x7FFFFFFF : *max_no_calls) : -1;
        cfg.matcher = kind == MatcherKind::PreCompute ? SGDX_MATCHER_PRECOMPUTE : SGDX_MATCHER_CACHED_HAMMING;
        cfg.index1_len = index_lengths ? (uint32_t)index_lengths->first : 0;
        cfg.index2_len = index_lengths ? (uint32_t)index_lengths->second : 0;
        cfg.device = device;
        cfg.slots = slots;
        cfg.kernel = SGDX_KERNEL_AUTO;
        check(sgdx_create(&cfg, reinterpret_cast<const uint8_t *>(table.data()), &h_), "sgdx_create");
    }
    GpuMatcher(const GpuMatcher &) = delete;
    GpuMatcher &operator=(const GpuMatcher &) = delete;
    ~GpuMatcher() override { sgdx_destroy(h_); }

    MatcherKind kind() const { return kind_; }
    size_t num_samples() const { return S_; }
    size_t barcode_len() const { return L_; }
    sgdx_handle *handle() const { return h_; }

    // trait Matcher: one barcode = one tiny device batch; use find_batch for throughput
    MatchResult find(std::string barcode) override {
        if (barcode.size() != L_) {
            // only reachable with --sample-barcode-in-fastq-header: every PreCompute key has length L (NoMatch);
            // the zip()-truncating Hamming case stays with the host's scalar code
            if (kind_ == MatcherKind::PreCompute) return MatchResult::NoMatch(std::move(barcode));
            throw std::invalid_argument("barcode length differs from the matcher's");
        }
        uint16_t idx = 0;
        uint8_t dist = 0;
        find_batch(reinterpret_cast<const uint8_t *>(barcode.data()), 1, &idx, &dist);
        if (idx == S_) return MatchResult::NoMatch(std::move(barcode));
        return MatchResult::Match(idx, dist, std::move(barcode));
    }
    // n barcodes of barcode_len() bytes each -> sample index (num_samples() = no match) and hamming distance
    void find_batch(const uint8_t *barcodes, uint64_t n, uint16_t *idx, uint8_t *dist, unsigned slot = 0) {
        check(sgdx_match_batch(h_, slot, barcodes, n, idx, dist), "sgdx_match_batch");
        check(sgdx_sync(h_, slot, nullptr), "sgdx_sync");
    }
    // the same through compact records (barcode_len() <= SGDX_COMPACT_MAX_LEN): packed on the host, 8 bytes per
    // read over PCIe; `records` is the caller's (ideally pinned) buffer of n words
    void find_batch_compact(const uint8_t *barcodes, uint64_t n, uint64_t *records, uint16_t *idx, uint8_t *dist,
                            unsigned slot = 0, unsigned pack_threads = 1) {
        check(sgdx_pack_compact_host(barcodes, n, (uint32_t)L_, records, pack_threads), "sgdx_pack_compact_host");
        check(sgdx_match_compact_batch(h_, slot, records, n, idx, dist), "sgdx_match_compact_batch");
        check(sgdx_sync(h_, slot, nullptr), "sgdx_sync");
    }
    // ... and through dense records (2 bits per base + an exception list for the reads with a no-call or a foreign
    // byte: SGDX_DENSE_BYTES(barcode_len()) bytes per read over PCIe).  The buffers are the library's vectors here;
    // a pipeline keeps pinned ones and calls sgdx_pack_dense_host / sgdx_match_dense_batch itself.
    void find_batch_dense(const uint8_t *barcodes, uint64_t n, uint16_t *idx, uint8_t *dist, unsigned slot = 0,
                          unsigned pack_threads = 1) {
        std::vector<uint8_t> dense(n * SGDX_DENSE_BYTES((uint32_t)L_) + 8);
        std::vector<uint32_t> ei(n ? n : 1);
        std::vector<uint64_t> er(n ? n : 1);
        uint64_t k = 0;
        check(sgdx_pack_dense_host(barcodes, n, (uint32_t)L_, dense.data(), ei.data(), er.data(), n, &k, pack_threads), "sgdx_pack_dense_host");
        check(sgdx_match_dense_batch(h_, slot, dense.data(), n, ei.data(), er.data(), k, idx, dist), "sgdx_match_dense_batch");
        check(sgdx_sync(h_, slot, nullptr), "sgdx_sync");
    }

  private:
    sgdx_handle *h_ = nullptr;
    MatcherKind kind_;
    size_t S_ = 0, L_ = 0;
};

struct CachedHammingDistanceMatcher : GpuMatcher {  // matcher.rs:80-109 (the LRU memo is a CPU detail)
    CachedHammingDistanceMatcher(const std::vector<SampleMetadata> &samples, size_t max_mismatches, size_t min_delta,
                                 size_t free_ns, std::optional<size_t> max_no_calls = {},
                                 std::optional<std::pair<size_t, size_t>> index_lengths = {}, int device = 0)
        : GpuMatcher(samples, MatcherKind::CachedHammingDistance, max_mismatches, min_delta, free_ns, max_no_calls,
                     index_lengths, device) {}
};
struct PreComputeMatcher : GpuMatcher {  // matcher.rs:121-262; free_ns is accepted and ignored, as in the reference
    PreComputeMatcher(const std::vector<SampleMetadata> &samples, size_t max_mismatches, size_t min_delta, size_t free_ns,
                      std::optional<size_t> max_no_calls = {},
                      std::optional<std::pair<size_t, size_t>> index_lengths = {}, int device = 0)
        : GpuMatcher(samples, MatcherKind::PreCompute, max_mismatches, min_delta, free_ns, max_no_calls, index_lengths,
                     device) {}
};

struct DemuxedGroupSampleMetrics {  // metrics.rs:407-437
    size_t perfect_matches = 0, one_mismatch_matches = 0, total_matches = 0;
    void update_with(const DemuxedGroupSampleMetrics &o) {
        perfect_matches += o.perfect_matches;
        one_mismatch_matches += o.one_mismatch_matches;
        total_matches += o.total_matches;
    }
};

struct SampleBarcodeHopTracker {  // metrics.rs:617-677
    size_t index1_length = 0, index2_length = 0;
    std::map<std::string, size_t> count_tracker;  // concatenated observed barcode -> count
    void update_with_self(const SampleBarcodeHopTracker &o) {
        for (const auto &kv : o.count_tracker) count_tracker[kv.first] += kv.second;
    }
};

struct DemuxedGroupMetrics {  // metrics.rs:297-328 (the fields this path fills)
    std::vector<DemuxedGroupSampleMetrics> per_sample_metrics;
    size_t total_templates = 0;
    std::optional<SampleBarcodeHopTracker> sample_barcode_hop_tracker;
    void update_with(const DemuxedGroupMetrics &o) {
        total_templates += o.total_templates;
        for (size_t i = 0; i < per_sample_metrics.size() && i < o.per_sample_metrics.size(); i++)
            per_sample_metrics[i].update_with(o.per_sample_metrics[i]);
        if (sample_barcode_hop_tracker && o.sample_barcode_hop_tracker)
            sample_barcode_hop_tracker->update_with_self(*o.sample_barcode_hop_tracker);
    }
};

struct BarcodeCount {  // metrics.rs:194-206
    std::string barcode;
    uint64_t count = 0;
};

struct BaseQualCounter {  // metrics.rs:468-497; the per-base update (499-518) runs on the device
    uint64_t q20_bases = 0, q30_bases = 0, index_bases_qual_sum = 0, index_bases_total_seen = 0, bases = 0;
    void update_with_self(const BaseQualCounter &o) {
        q20_bases += o.q20_bases;
        q30_bases += o.q30_bases;
        index_bases_qual_sum += o.index_bases_qual_sum;
        index_bases_total_seen += o.index_bases_total_seen;
        bases += o.bases;
    }
};

// `--read-structures` strings (opts.rs:84-92): <length><kind> segments, kinds T B M S, `+` only on the last one
struct ReadStructure {
    std::vector<sgdx_segment> segments;  // source is filled in by the calls that take several structures
    static ReadStructure from_str(const std::string &text) {
        ReadStructure rs;
        uint32_t off = 0;
        size_t i = 0;
        while (i < text.size()) {
            uint32_t len = 0;
            bool plus = false;
            if (text[i] == '+') {
                plus = true;
                i++;
            } else {
                size_t j = i;
                while (j < text.size() && text[j] >= '0' && text[j] <= '9') len = len * 10 + (uint32_t)(text[j++] - '0');
                if (j == i || len == 0) throw std::invalid_argument("read structure: expected a positive length in " + text);
                i = j;
            }
            if (i >= text.size()) throw std::invalid_argument("read structure: missing segment kind in " + text);
            const char k = (char)(text[i] >= 'a' ? text[i] - 32 : text[i]);
            i++;
            if (k != 'T' && k != 'B' && k != 'M' && k != 'S') throw std::invalid_argument("read structure: unknown kind in " + text);
            if (plus && i != text.size()) throw std::invalid_argument("read structure: only the last segment may be `+`: " + text);
            rs.segments.push_back(sgdx_segment{0u, off, plus ? SGDX_SEG_TO_END : len, (uint32_t)k});
            off += len;
        }
        if (rs.segments.empty()) throw std::invalid_argument("empty read structure");
        return rs;
    }
};

struct DemuxedGroup {  // demux.rs:637-645, minus the routed reads
    std::vector<uint16_t> sample_indices;  // undetermined index for NoMatch
    std::vector<uint8_t> hamming_dists;    // SGDX_NO_DIST for NoMatch
};

// The assignment + counting step of Demultiplexer::demultiplex, batched.  `samples` holds ALL samples, the last one
// is Undetermined (demux.rs:221,259); the matcher sees samples[..n-1] (run.rs:201,220).
class BarcodeAssignmentStage {
  public:
    BarcodeAssignmentStage(const std::vector<SampleMetadata> &samples, MatcherKind kind, size_t max_mismatches,
                           size_t min_delta, size_t free_ns, std::optional<size_t> max_no_calls = {},
                           std::optional<std::pair<size_t, size_t>> index_lengths = {}, int device = 0)
        : undetermined_(samples.size() - 1), index_lengths_(index_lengths),
          matcher_(std::vector<SampleMetadata>(samples.begin(), samples.end() - 1), kind, max_mismatches, min_delta,
                   free_ns, max_no_calls, index_lengths, device) {}

    size_t get_undetermined_index() const { return undetermined_; }
    GpuMatcher &matcher() { return matcher_; }

    DemuxedGroup demultiplex(const std::vector<std::string> &barcodes) {
        std::string flat;
        for (const auto &b : barcodes) {
            if (b.size() != matcher_.barcode_len()) throw std::invalid_argument("device batches need barcodes of length L");
            flat += b;
        }
        DemuxedGroup g;
        g.sample_indices.resize(barcodes.size());
        g.hamming_dists.resize(barcodes.size());
        matcher_.find_batch(reinterpret_cast<const uint8_t *>(flat.data()), barcodes.size(), g.sample_indices.data(),
                            g.hamming_dists.data());
        return g;
    }

    // UnmatchedCounterThread::new (run.rs:185-190): from now on every batch also counts its NoMatch barcodes on the device
    void collect_unmatched(uint64_t max_map_size = 5000000, uint64_t downsize_to = 5000) {
        check(sgdx_unmatched_enable(matcher_.handle(), max_map_size, downsize_to), "sgdx_unmatched_enable");
    }
    // Lengths of the sample-barcode (B) segments in FASTQ order, then segment order: the delimited barcode has a `+`
    // between every two of them (ExtractedBarcode::add_bases, demux.rs:56-62).  Default: one segment per barcode read.
    void set_barcode_segment_lengths(std::vector<size_t> lengths) { segment_lengths_ = std::move(lengths); }
    std::string delimit(const std::string &concatenated) const {
        std::vector<size_t> cuts = segment_lengths_;
        if (cuts.empty() && index_lengths_) cuts = {index_lengths_->first};
        std::string out;
        size_t pos = 0;
        for (size_t ln : cuts) {
            if (pos >= concatenated.size()) break;
            if (pos) out += '+';
            out += concatenated.substr(pos, ln);
            pos += ln;
        }
        if (pos < concatenated.size()) {
            if (pos) out += '+';
            out += concatenated.substr(pos);
        }
        return out;
    }
    // rows of most_frequent_unmatched.tsv (metrics.rs:177-190): delimited barcode (demux.rs:56-62,504-522), descending count
    std::vector<BarcodeCount> most_frequent_unmatched(uint64_t n) {
        const size_t L = matcher_.barcode_len();
        std::vector<uint8_t> keys((n ? n : 1) * L);
        std::vector<uint64_t> counts(n ? n : 1);
        uint64_t w = 0;
        check(sgdx_unmatched_top(matcher_.handle(), n, keys.data(), counts.data(), &w), "sgdx_unmatched_top");
        std::vector<BarcodeCount> rows;
        for (uint64_t i = 0; i < w; i++) {
            std::string k(reinterpret_cast<const char *>(keys.data()) + i * L, L);
            rows.push_back({delimit(k), counts[i]});
        }
        return rows;
    }

    // BaseQualCounter::update over one FASTQ's quality strings of a batch that is resident on the device, keyed
    // by the batch's sample-index array (demux.rs:582-583); d_lens may be null (all reads `stride` long)
    void count_base_qualities(const uint8_t *d_quals, uint64_t n, uint32_t stride, const ReadStructure &rs,
                              const uint16_t *d_sample_indices, const uint32_t *d_lens = nullptr, unsigned slot = 0) {
        check(sgdx_qual_counts_device(matcher_.handle(), slot, d_quals, n, stride, d_lens, rs.segments.data(),
                                      (uint32_t)rs.segments.size(), d_sample_indices), "sgdx_qual_counts_device");
    }
    // Demultiplex::demultiplex (demux.rs:164-173,445-584) for a device-resident batch: extraction, assignment with
    // its counters (and the unmatched table when collect_unmatched() was called), quality counters of every FASTQ
    // that brings qualities, routing when out.d_order is set -- all queued on the slot's stream
    void demultiplex_device(const std::vector<sgdx_fastq_view> &fastqs, const std::vector<ReadStructure> &structures,
                            uint64_t n, const sgdx_demux_out &out, unsigned slot = 0) {
        if (fastqs.size() != structures.size()) throw std::invalid_argument("one read structure per FASTQ (run.rs:109-119)");
        std::vector<sgdx_segment> segs;
        for (size_t f = 0; f < structures.size(); f++)
            for (sgdx_segment g : structures[f].segments) {
                g.source = (uint32_t)f;
                segs.push_back(g);
            }
        check(sgdx_demultiplex_device(matcher_.handle(), slot, fastqs.data(), (uint32_t)fastqs.size(), segs.data(),
                                      (uint32_t)segs.size(), n, &out), "sgdx_demultiplex_device");
    }
    std::vector<BaseQualCounter> base_qual_counters() {  // one per sample, Undetermined last
        std::vector<sgdx_base_qual> raw(matcher_.num_samples() + 1);
        check(sgdx_base_qual_counters(matcher_.handle(), raw.data(), nullptr), "sgdx_base_qual_counters");
        std::vector<BaseQualCounter> out;
        for (const auto &r : raw)
            out.push_back({r.q20_bases, r.q30_bases, r.index_bases_qual_sum, r.index_bases_total_seen, r.bases});
        return out;
    }

    // cumulative counters of every batch this stage has processed (device resident, read on demand)
    DemuxedGroupMetrics metrics() {
        const size_t S = matcher_.num_samples();
        std::vector<uint64_t> per((S + 1) * 3);
        sgdx_totals t{};
        check(sgdx_counters(matcher_.handle(), per.data(), &t), "sgdx_counters");
        DemuxedGroupMetrics m;
        m.total_templates = t.total_templates;
        for (size_t s = 0; s <= S; s++) m.per_sample_metrics.push_back({per[s * 3 + 1], per[s * 3 + 2], per[s * 3]});
        if (index_lengths_) {
            SampleBarcodeHopTracker tr{index_lengths_->first, index_lengths_->second, {}};
            uint64_t n = 0, w = 0;
            check(sgdx_hop_count(matcher_.handle(), &n), "sgdx_hop_count");
            const size_t L = matcher_.barcode_len();
            std::vector<uint8_t> keys((n ? n : 1) * L);
            std::vector<uint64_t> counts(n ? n : 1);
            check(sgdx_hop_export(matcher_.handle(), keys.data(), counts.data(), n ? n : 1, &w), "sgdx_hop_export");
            for (uint64_t i = 0; i < w; i++)
                tr.count_tracker[std::string(reinterpret_cast<const char *>(keys.data()) + i * L, L)] = counts[i];
            m.sample_barcode_hop_tracker = std::move(tr);
        }
        return m;
    }

  private:
    size_t undetermined_;
    std::optional<std::pair<size_t, size_t>> index_lengths_;
    std::vector<size_t> segment_lengths_;  // B segments, FASTQ order then segment order; empty = one per barcode read
    GpuMatcher matcher_;
};

}  // namespace sgdx
