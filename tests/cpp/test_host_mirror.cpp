// test_host_mirror.cpp -- the reference's own known-answer tests for the assignment path, written against the
// C++ host mirror (singular-demux_b200/host/sgdx_host.hpp) so that they read like the originals.  Needs a CUDA
// device (run by tests/test_gpu_parity.py::test_cpp_host_mirror_known_answers on a B200).
//   matcher.rs:550-622   test_find_two_samples / matcher.rs:682-727 readme examples
//   demux.rs:1181-1267   delta too small, too many mismatches, too many Ns
//   run.rs:593-690       test_end_to_end_simple (assignment + per-sample metrics + total_templates)
//   run.rs:1184-1238     test_end_to_end_with_all_matchers
//   run.rs:1382-1535     dual-index hop metrics / sample metrics, run.rs:1539-1605 single index no match
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../singular-demux_b200/host/sgdx_host.hpp"

using namespace sgdx;

static int failures = 0;
#define EXPECT(cond)                                                                   \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
            failures++;                                                                \
        }                                                                              \
    } while (0)

static std::vector<SampleMetadata> samples_of(const std::vector<std::string> &barcodes, bool with_undetermined) {
    std::vector<SampleMetadata> v;
    for (size_t i = 0; i < barcodes.size(); i++) v.push_back({"Sample" + std::to_string(i + 1), barcodes[i], i});
    if (with_undetermined) v.push_back({UNDETERMINED_NAME, std::string(barcodes[0].size(), 'N'), barcodes.size()});
    return v;
}
static const MatcherKind kBoth[] = {MatcherKind::CachedHammingDistance, MatcherKind::PreCompute};

// utils.rs:338-341
static const std::vector<std::string> kFour = {"AAAAAAAAGATTACAGA", "CCCCCCCCGATTACAGA", "GGGGGGGGGATTACAGA",
                                               "GGGGGGTTGATTACAGA"};

static void test_find_two_samples() {  // matcher.rs:550-622
    for (MatcherKind k : kBoth) {
        GpuMatcher m(samples_of({"AAAA", "AACC"}, false), k, 2, 1, 0);
        MatchResult r = m.find("AAAA");
        EXPECT(r.is_match() && r.sample_index == 0 && r.hamming_dist == 0 && r.barcode == "AAAA");
        EXPECT(m.find("AAAG").is_no_match());  // distances 1 and 2: delta 1 <= min_delta 1
    }
}

static void test_record_forms_agree() {  // the same batch as ASCII, compact records and dense records + exceptions
    for (MatcherKind k : kBoth) {
        GpuMatcher m(samples_of({"AAAAAAAA", "CCCCGGGG", "ACGTACGT"}, false), k, 1, 1, 1);
        const std::string reads = "AAAAAAAA" "AAAAAAAT" "CCCCGGGG" "NCCCGGGG" "ACGTACGA" "TTTTTTTT" "ACGTNNGT" "RAAAAAAA" "CCCCGGGC";
        const uint64_t n = reads.size() / 8;
        std::vector<uint16_t> a(n), b(n), c(n);
        std::vector<uint8_t> da(n), db(n), dc(n);
        std::vector<uint64_t> rec(n);
        const uint8_t *p = reinterpret_cast<const uint8_t *>(reads.data());
        m.find_batch(p, n, a.data(), da.data());
        m.find_batch_compact(p, n, rec.data(), b.data(), db.data());
        m.find_batch_dense(p, n, c.data(), dc.data());
        EXPECT(a == b && a == c && da == db && da == dc);
        EXPECT(a[0] == 0 && da[0] == 0 && a[1] == 0 && da[1] == 1 && a[5] == 3);
    }
}

static void test_readme_examples() {  // matcher.rs:682-727: max_mismatches 3, min_delta 1
    struct Case { std::vector<std::string> samples; std::string observed; bool match; size_t dist; };
    const Case cases[] = {
        {{"AAA", "TTT"}, "GGA", false, 0},        // best 2, next 3
        {{"AAA", "TTT"}, "GAA", true, 1},         // best 1, next 3
        {{"AA", "TT"}, "AA", true, 0},            // best 0, next 2
        {{"A", "T"}, "A", false, 0},              // best 0, next 1
        {{"AAAAAA", "TTTTTT"}, "GGGAAA", true, 3},   // best 3, next 6
        {{"AAAAAA", "TTTTTT"}, "GGGGAA", false, 0},  // best 4 > 3
        {{"AA", "TT"}, "GG", false, 0},           // best 2, next 2
    };
    for (MatcherKind k : kBoth)
        for (const Case &c : cases) {
            GpuMatcher m(samples_of(c.samples, false), k, 3, 1, 0);
            MatchResult r = m.find(c.observed);
            EXPECT(r.is_match() == c.match);
            if (c.match) EXPECT(r.sample_index == 0 && r.hamming_dist == c.dist);
        }
}

static void test_demux_known_answers() {
    {   // demux.rs:1181-1204: exact SAMPLE_BARCODE_4, min_delta 3 -> Undetermined (delta to sample 3 is 2)
        BarcodeAssignmentStage st(samples_of(kFour, true), MatcherKind::CachedHammingDistance, 2, 3, 2, 1);
        EXPECT(st.demultiplex({"GGGGGGTTGATTACAGA"}).sample_indices[0] == st.get_undetermined_index());
    }
    for (MatcherKind k : kBoth) {  // demux.rs:1208-1235: one mismatch, allowed 0
        BarcodeAssignmentStage st(samples_of(kFour, true), k, 0, 1, 2, 1);
        EXPECT(st.demultiplex({"AAAAAAAAGATTACAGT"}).sample_indices[0] == 4);
    }
    for (MatcherKind k : kBoth) {  // demux.rs:1239-1267: one N, max_no_calls 0
        BarcodeAssignmentStage st(samples_of(kFour, true), k, 1, 1, 2, 0);
        EXPECT(st.demultiplex({"NAAAAAAAGATTACAGA"}).sample_indices[0] == 4);
    }
}

static void test_end_to_end_simple() {  // run.rs:593-690 with the defaults of opts.rs:535-560
    BarcodeAssignmentStage st(samples_of(kFour, true), MatcherKind::CachedHammingDistance, 2, 1, 1, 1);
    DemuxedGroup g = st.demultiplex({"AAAAAAAAGATTACAGA", "AAAAAAAAGATTACAGT", "AAAAAAAAGATTACTTT", "GGGGGGTCGATTACAGA",
                                     "AAAAAAAAGANNNNNNN"});
    const uint16_t want_idx[] = {0, 0, 4, 4, 4};
    for (int i = 0; i < 5; i++) EXPECT(g.sample_indices[i] == want_idx[i]);
    EXPECT(g.hamming_dists[0] == 0 && g.hamming_dists[1] == 1 && g.hamming_dists[2] == SGDX_NO_DIST);
    DemuxedGroupMetrics m = st.metrics();
    EXPECT(m.total_templates == 5);
    EXPECT(m.per_sample_metrics[0].total_matches == 2 && m.per_sample_metrics[0].perfect_matches == 1 &&
           m.per_sample_metrics[0].one_mismatch_matches == 1);
    EXPECT(m.per_sample_metrics[4].total_matches == 3 && m.per_sample_metrics[4].perfect_matches == 0);
    EXPECT(m.per_sample_metrics[1].total_matches == 0 && m.per_sample_metrics[2].total_matches == 0 &&
           m.per_sample_metrics[3].total_matches == 0);
}

static void test_end_to_end_with_all_matchers() {  // run.rs:1184-1238: 8 bp, M=1, delta=1
    for (MatcherKind k : kBoth) {
        BarcodeAssignmentStage st(samples_of({"AAAAAAAA", "ACGTACGT"}, true), k, 1, 1, 1, 1);
        DemuxedGroup g = st.demultiplex({"AAAAAAAA", "AAAATAAA", "ACAAATAA"});
        EXPECT(g.sample_indices[0] == 0 && g.sample_indices[1] == 0 && g.sample_indices[2] == 2);
        DemuxedGroupMetrics m = st.metrics();
        EXPECT(m.per_sample_metrics[0].total_matches == 2 && m.per_sample_metrics[1].total_matches == 0 &&
               m.per_sample_metrics[2].total_matches == 1);
    }
    EXPECT(select_matcher(8) == MatcherKind::PreCompute && select_matcher(17) == MatcherKind::CachedHammingDistance);
    EXPECT(select_matcher(17, MatcherKind::PreCompute) == MatcherKind::PreCompute);
    EXPECT(select_matcher(8, MatcherKind::CachedHammingDistance) == MatcherKind::PreCompute);  // run.rs:196
}

static void test_dual_index_metrics() {
    {   // run.rs:1382-1451: TTT+GGG, CCC+AAA; observed TTT+AAA hops
        BarcodeAssignmentStage st(samples_of({"TTTGGG", "CCCAAA"}, true), select_matcher(6), 2, 1, 1, 1,
                                  std::make_pair<size_t, size_t>(3, 3));
        EXPECT(st.demultiplex({"TTTAAA"}).sample_indices[0] == 2);
        DemuxedGroupMetrics m = st.metrics();
        EXPECT(m.sample_barcode_hop_tracker && m.sample_barcode_hop_tracker->count_tracker.size() == 1 &&
               m.sample_barcode_hop_tracker->count_tracker["TTTAAA"] == 1);
    }
    {   // run.rs:1455-1535: TTT+AAA is a sample
        BarcodeAssignmentStage st(samples_of({"TTTAAA", "CCCGGG"}, true), select_matcher(6), 2, 1, 1, 1,
                                  std::make_pair<size_t, size_t>(3, 3));
        EXPECT(st.demultiplex({"TTTAAA"}).sample_indices[0] == 0);
        DemuxedGroupMetrics m = st.metrics();
        EXPECT(m.per_sample_metrics[0].total_matches == 1 && m.sample_barcode_hop_tracker->count_tracker.empty());
    }
    {   // run.rs:1539-1605: single index, no match, no hop tracker
        BarcodeAssignmentStage st(samples_of({"CCC"}, true), select_matcher(3), 2, 1, 1, 1);
        EXPECT(st.demultiplex({"TTT"}).sample_indices[0] == 1);
        DemuxedGroupMetrics m = st.metrics();
        EXPECT(m.per_sample_metrics[1].total_matches == 1 && !m.sample_barcode_hop_tracker);
    }
}

static void test_metrics_merge_and_errors() {
    // metrics.rs:315-328: merging is a plain sum -> two stages over halves equal one stage over everything
    std::vector<std::string> reads = {"AAAAAAAAGATTACAGA", "AAAAAAAAGATTACAGT", "CCCCCCCCGATTACAGA", "GGGGGGTCGATTACAGA"};
    BarcodeAssignmentStage whole(samples_of(kFour, true), MatcherKind::CachedHammingDistance, 2, 1, 1, 1);
    whole.demultiplex(reads);
    BarcodeAssignmentStage a(samples_of(kFour, true), MatcherKind::CachedHammingDistance, 2, 1, 1, 1);
    BarcodeAssignmentStage b(samples_of(kFour, true), MatcherKind::CachedHammingDistance, 2, 1, 1, 1);
    a.demultiplex({reads[0], reads[1]});
    b.demultiplex({reads[2], reads[3]});
    DemuxedGroupMetrics m = a.metrics(), w = whole.metrics();
    m.update_with(b.metrics());
    EXPECT(m.total_templates == w.total_templates);
    for (size_t i = 0; i < w.per_sample_metrics.size(); i++)
        EXPECT(m.per_sample_metrics[i].total_matches == w.per_sample_metrics[i].total_matches &&
               m.per_sample_metrics[i].perfect_matches == w.per_sample_metrics[i].perfect_matches);
    bool threw = false;
    try {
        GpuMatcher bad(samples_of({"ACGN"}, false), MatcherKind::CachedHammingDistance, 1, 1, 1);  // N in a sample barcode
    } catch (const Error &e) {
        threw = e.code == SGDX_EINVAL;
    }
    EXPECT(threw);
}

static void test_unmatched_and_read_structures() {
    // run.rs:652-689: the three unmatched barcodes of the simple end-to-end run, once each
    BarcodeAssignmentStage st(samples_of(kFour, true),
                              MatcherKind::CachedHammingDistance, 2, 1, 1, 1);
    st.collect_unmatched();
    st.demultiplex({"AAAAAAAAGATTACAGA", "AAAAAAAAGATTACAGT", "AAAAAAAAGATTACTTT", "GGGGGGTCGATTACAGA", "AAAAAAAAGANNNNNNN"});
    auto rows = st.most_frequent_unmatched(1000);
    EXPECT(rows.size() == 3);
    for (const auto &r : rows) EXPECT(r.count == 1);
    EXPECT(rows[0].barcode == "AAAAAAAAGANNNNNNN" && rows[1].barcode == "AAAAAAAAGATTACTTT" && rows[2].barcode == "GGGGGGTCGATTACAGA");
    // run.rs:1444-1451: the hopped barcode, delimited
    BarcodeAssignmentStage dual(samples_of({"TTTGGG", "CCCAAA"}, true), MatcherKind::PreCompute, 2, 1, 1, 1,
                                std::make_pair<size_t, size_t>(3, 3));
    dual.collect_unmatched();
    dual.demultiplex({"TTTAAA"});
    auto hop = dual.most_frequent_unmatched(10);
    EXPECT(hop.size() == 1 && hop[0].barcode == "TTT+AAA" && hop[0].count == 1);
    // demux.rs:56-62: a `+` between every two sample-barcode segments -- 2B1S1B + 2B here, three pieces
    dual.set_barcode_segment_lengths({2, 1, 3});
    EXPECT(dual.delimit("TTTAAA") == "TT+T+AAA" && dual.most_frequent_unmatched(10)[0].barcode == "TT+T+AAA");
    EXPECT(st.delimit("AAAAAAAAGATTACAGA") == "AAAAAAAAGATTACAGA");  // one barcode read, one segment: no delimiter
    // read structures (opts.rs:84-92)
    auto rs = ReadStructure::from_str("8B+T");
    EXPECT(rs.segments.size() == 2 && rs.segments[0].length == 8 && rs.segments[0].kind == 'B');
    EXPECT(rs.segments[1].offset == 8 && rs.segments[1].length == SGDX_SEG_TO_END && rs.segments[1].kind == 'T');
    bool threw = false;
    try {
        ReadStructure::from_str("+T8B");
    } catch (const std::invalid_argument &) {
        threw = true;
    }
    EXPECT(threw);
    BaseQualCounter a{1, 2, 3, 4, 5}, b{10, 20, 30, 40, 50};
    a.update_with_self(b);  // metrics.rs:489-496
    EXPECT(a.q20_bases == 11 && a.q30_bases == 22 && a.index_bases_qual_sum == 33 && a.index_bases_total_seen == 44 && a.bases == 55);
    auto q = st.base_qual_counters();
    EXPECT(q.size() == 5 && q[0].bases == 0);
}

int main() {
    try {
        test_find_two_samples();
        test_record_forms_agree();
        test_readme_examples();
        test_demux_known_answers();
        test_end_to_end_simple();
        test_end_to_end_with_all_matchers();
        test_dual_index_metrics();
        test_metrics_merge_and_errors();
        test_unmatched_and_read_structures();
    } catch (const Error &e) {  // e.g. no CUDA device: there is no CPU fallback to run these on
        std::fprintf(stderr, "sgdx::Error (code %d): %s\n", e.code, e.what());
        return 2;
    }
    if (failures) {
        std::fprintf(stderr, "%d expectation(s) failed\n", failures);
        return 1;
    }
    std::puts("host mirror: all reference known answers reproduced");
    return 0;
}
