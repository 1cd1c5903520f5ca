#!/usr/bin/env python3
"""Writes tests/golden/reference_known_answers.json.

Every entry is TRANSCRIBED BY HAND from an assertion in the reference's own test-suite
(/root/reference/src/lib/<file>:<lines>, cited per entry); nothing here is computed by our oracle
or by our kernels.  The reference is Rust and cannot run in this image, so these known answers are
what pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

Conventions: matcher 0 = CachedHammingDistance, 1 = PreCompute; max_no_calls -1 = None;
expected sample index == len(samples) means Undetermined (demux.rs:259,535-538).
"""
import json
import os

S1, S2, S3, S4 = "AAAAAAAAGATTACAGA", "CCCCCCCCGATTACAGA", "GGGGGGGGGATTACAGA", "GGGGGGTTGATTACAGA"  # utils.rs:338-341
PRESET = [S1, S2, S3, S4]

golden = {
    "hamming_distance": [  # matcher.rs:454-497  [alpha, beta, free_ns, expected]
        ["GATTACA", "GATTACA", 0, 0],
        ["GATTACA", "GACCACA", 0, 2],
        ["GATTACA", "GANNACA", 0, 2],
        ["GATTACA", "CTAATGT", 0, 7],
        ["GATTACN", "GATTACN", 0, 1],
        ["GATTACA", "GATTACN", 0, 1],
        ["GATTACN", "GATTACA", 0, 1],
        ["GATTACN", "GATTACN", 1, 0],
        ["GATTACA", "GATTACN", 1, 0],
        ["GATTACN", "GATTACA", 1, 0],
    ],
    "should_reject_delta": [  # matcher.rs:534-548  [delta, min_delta, expected]
        [0, 0, True], [1, 0, False], [2, 0, False], [1, 3, True], [3, 3, True], [4, 3, False],
    ],
    "precompute_map": {  # matcher.rs:412-447: exact map for AAA/TTT, M=1, delta=2
        "samples": ["AAA", "TTT"], "max_mismatches": 1, "min_delta": 2,
        "expected": {
            "AAG": [0, 1], "AGA": [0, 1], "GAA": [0, 1], "AAC": [0, 1], "ACA": [0, 1], "CAA": [0, 1],
            "AAN": [0, 1], "ANA": [0, 1], "NAA": [0, 1], "AAA": [0, 0],
            "TTG": [1, 1], "TGT": [1, 1], "GTT": [1, 1], "TTC": [1, 1], "TCT": [1, 1], "CTT": [1, 1],
            "TTN": [1, 1], "TNT": [1, 1], "NTT": [1, 1], "TTT": [1, 0],
        },
        # matcher.rs:448-451: build_map(samples, 3, 0) == build_map(samples, 10, 0)
        "degenerate_equal": [[3, 0], [10, 0]],
    },
    # find(): each case is checked with BOTH matchers unless "matchers" says otherwise.
    # query = [observed, is_match, sample_index_or_null, dist_or_null]
    "matcher_cases": [
        {"src": "matcher.rs:550-622", "samples": ["AAAA", "AACC"], "max_mismatches": 2, "min_delta": 1,
         "free_ns": 0, "matchers": [0, 1],
         "queries": [["AAAA", True, 0, 0], ["AAAG", False, None, None]]},
    ],
    # demux-level cases (N gate + matcher + index + counters + hop + unmatched)
    "demux_cases": [
        {"src": "run.rs:593-690 test_end_to_end_simple (fixture run.rs:535-578, defaults opts.rs:535-560)",
         "samples": PRESET, "max_mismatches": 2, "min_delta": 1, "free_ns": 1, "max_no_calls": 1,
         "matchers": [0], "index1_len": 0, "index2_len": 0,
         "reads": [S1, "AAAAAAAAGATTACAGT", "AAAAAAAAGATTACTTT", "GGGGGGTCGATTACAGA", "AAAAAAAAGANNNNNNN"],
         "expected_idx": [0, 0, 4, 4, 4],
         "expected_dist": [0, 1, None, None, None],
         # [total_matches, perfect_matches, one_mismatch_matches] per sample, Undetermined last
         "per_sample": [[2, 1, 1], [0, 0, 0], [0, 0, 0], [0, 0, 0], [3, 0, 0]],
         "total_templates": 5,
         "unmatched": {"AAAAAAAAGATTACTTT": 1, "GGGGGGTCGATTACAGA": 1, "AAAAAAAAGANNNNNNN": 1}},
        {"src": "demux.rs:1181-1204 (matcher free_ns=2 per demux.rs:748-753)",
         "samples": PRESET, "max_mismatches": 2, "min_delta": 3, "free_ns": 2, "max_no_calls": 1,
         "matchers": [0], "index1_len": 0, "index2_len": 0,
         "reads": [S4], "expected_idx": [4], "expected_dist": [None]},
        {"src": "demux.rs:1208-1235", "samples": PRESET, "max_mismatches": 0, "min_delta": 1, "free_ns": 2,
         "max_no_calls": 1, "matchers": [0, 1], "index1_len": 0, "index2_len": 0,
         "reads": ["AAAAAAAAGATTACAGT"], "expected_idx": [4], "expected_dist": [None]},
        {"src": "demux.rs:1239-1267 (N gate, max_no_calls=0)", "samples": PRESET, "max_mismatches": 1,
         "min_delta": 1, "free_ns": 2, "max_no_calls": 0, "matchers": [0, 1], "index1_len": 0,
         "index2_len": 0, "reads": ["NAAAAAAAGATTACAGA"], "expected_idx": [4], "expected_dist": [None]},
        {"src": "run.rs:1184-1238 test_end_to_end_with_all_matchers (None -> PreCompute by run.rs:196)",
         "samples": ["AAAAAAAA", "ACGTACGT"], "max_mismatches": 1, "min_delta": 1, "free_ns": 1,
         "max_no_calls": 1, "matchers": [0, 1], "index1_len": 0, "index2_len": 0,
         "reads": ["AAAAAAAA", "AAAATAAA", "ACAAATAA"], "expected_idx": [0, 0, 2],
         "per_sample_totals": [2, 0, 1]},
        {"src": "run.rs:1382-1451 test_dual_index_hop_metrics (6 bp -> PreCompute; defaults M=2 d=1)",
         "samples": ["TTTGGG", "CCCAAA"], "max_mismatches": 2, "min_delta": 1, "free_ns": 1,
         "max_no_calls": 1, "matchers": [1], "index1_len": 3, "index2_len": 3,
         "reads": ["TTTAAA"], "expected_idx": [2], "hops": {"TTTAAA": 1}, "unmatched": {"TTTAAA": 1}},
        {"src": "run.rs:1455-1535 test_dual_index_sample_metrics", "samples": ["TTTAAA", "CCCGGG"],
         "max_mismatches": 2, "min_delta": 1, "free_ns": 1, "max_no_calls": 1, "matchers": [1],
         "index1_len": 3, "index2_len": 3, "reads": ["TTTAAA"], "expected_idx": [0],
         "per_sample_totals": [1, 0, 0], "hops": {}, "unmatched": {}},
        {"src": "run.rs:1539-1605 test_single_index_nomatch_sample_metrics", "samples": ["CCC"],
         "max_mismatches": 2, "min_delta": 1, "free_ns": 1, "max_no_calls": 1, "matchers": [1],
         "index1_len": 0, "index2_len": 0, "reads": ["TTT"], "expected_idx": [1],
         "per_sample_totals": [0, 1], "unmatched": {"TTT": 1}},
        # 1000 + 1000 exact reads of two samples through `+T +T 8B` (one barcode read: no hop checker).  The test
        # builds its PreComputeMatcher over all three metadata entries, Undetermined's all-N barcode included
        # (demux.rs:1684); production hands the matcher samples[0..n-1] (run.rs:201,220) and so does this case --
        # an exact read is Match(sample, 0) either way.
        {"src": "demux.rs:1583-1730 test_demux_simple_multi_read", "samples": ["ACTGACTG", "GTCAGTCA"],
         "max_mismatches": 1, "min_delta": 0, "free_ns": 2, "max_no_calls": -1, "matchers": [1],
         "index1_len": 0, "index2_len": 0, "reads": ["ACTGACTG"] * 1000 + ["GTCAGTCA"] * 1000,
         "expected_idx": [0] * 1000 + [1] * 1000, "expected_dist": [0] * 2000,
         "per_sample": [[1000, 1000, 0], [1000, 1000, 0], [0, 0, 0]], "total_templates": 2000, "unmatched": {}},
        # --sample-barcode-in-fastq-header: the barcode comes from the header, single index (no delimiter) ...
        {"src": "demux.rs:1758-1777 test_demux_fragment_reads_sample_barcode_from_fastq_header_single_index "
                "(DemuxTestContext::demux_structures: M=2 d=1 max_no_calls=1, matcher free_ns=2)",
         "samples": PRESET, "max_mismatches": 2, "min_delta": 1, "free_ns": 2, "max_no_calls": 1, "matchers": [0],
         "index1_len": 0, "index2_len": 0, "reads": [S1], "expected_idx": [0], "expected_dist": [0],
         "per_sample_totals": [1, 0, 0, 0, 0]},
        # ... and dual index: `AAAAAAAAGA+TTACAGA` in the header; ExtractedBarcode::from_delimited (demux.rs:42-53)
        # drops the first `+` for the matcher and keeps the delimited form for the output header
        {"src": "demux.rs:1779-1799 test_demux_fragment_reads_sample_barcode_from_fastq_header_dual_index",
         "samples": PRESET, "max_mismatches": 2, "min_delta": 1, "free_ns": 2, "max_no_calls": 1, "matchers": [0],
         "index1_len": 0, "index2_len": 0, "header_barcodes": [S1[:10] + "+" + S1[10:]], "reads": [S1],
         "expected_idx": [0], "expected_dist": [0], "per_sample_totals": [1, 0, 0, 0, 0]},
    ],
}

# matcher.rs:625-727 README examples: M=3, delta=1; samples 'A'*next and 'T'*next; observed =
# 'G'*best + 'A'*(next-best); both matchers must agree with should_match.
for should_match, best, nxt in [(False, 2, 3), (True, 1, 3), (True, 0, 2), (False, 0, 1), (True, 3, 6),
                                (False, 4, 6), (False, 2, 2)]:
    golden["matcher_cases"].append({
        "src": f"matcher.rs:682-727 readme example best={best} next={nxt}",
        "samples": ["A" * nxt, "T" * nxt], "max_mismatches": 3, "min_delta": 1, "free_ns": 0,
        "matchers": [0, 1],
        "queries": [["G" * best + "A" * (nxt - best), should_match, 0 if should_match else None,
                     best if should_match else None]],
    })

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_known_answers.json")
    with open(out, "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote", out)
