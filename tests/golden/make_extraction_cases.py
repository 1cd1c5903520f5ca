#!/usr/bin/env python3
"""Transcribes the reference's segment-extraction tests (/root/reference/src/lib/demux.rs:867-1053, test context
demux.rs:797-803: allowed_mismatches 2, min_delta 1, max_no_calls Some(1), free_ns 2, samples utils.rs:338-341) into
tests/golden/reference_extraction_cases.json.  Every case is one template; the reference asserts that it lands in
sample 1 with the listed template bases and that the header carries the delimited sample barcode.  Default
qualities of the test records are '!' (q = 0), so q20 = q30 = 0 and `bases` = the template length.

Run: python tests/golden/make_extraction_cases.py   (no reference import: the values are the tests' own literals)"""
import json
import os

S1 = "AAAAAAAAGATTACAGA"  # SAMPLE_BARCODE_1, utils.rs:338
SAMPLES = [S1, "CCCCCCCCGATTACAGA", "GGGGGGGGGATTACAGA", "GGGGGGTTGATTACAGA"]

CASES = [
    {"src": "demux.rs:871-891 test_demux_fragment_reads", "structures": ["17B100T"], "bases": [S1 + "A" * 100],
     "delimited": S1, "template_bases": 100},
    {"src": "demux.rs:895-928 test_demux_paired_reads_inline_sample_barcodes", "structures": ["8B100T", "9B100T"],
     "bases": [S1[:8] + "A" * 100, S1[8:] + "T" * 100], "delimited": S1[:8] + "+" + S1[8:], "template_bases": 200},
    {"src": "demux.rs:932-969 test_demux_dual_indexed_paired_end_reads", "structures": ["8B", "100T", "9B", "100T"],
     "bases": [S1[:8], "A" * 100, S1[8:], "T" * 100], "delimited": S1[:8] + "+" + S1[8:], "template_bases": 200},
    {"src": "demux.rs:973-1011 test_demux_very_weird_set_of_reads",
     "structures": ["4B4M8S", "4B100T", "100S3B", "6B1S1M1T"],
     "bases": [S1[:4] + "CCCCGGGGTTTT", S1[4:8] + "A" * 100, "T" * 100 + S1[8:11], S1[11:] + "AAT"],
     "delimited": "+".join([S1[:4], S1[4:8], S1[8:11], S1[11:]]), "template_bases": 101},
    {"src": "demux.rs:1015-1043 test_demux_a_read_structure_with_multiple_independent_template_segments_in_one_read",
     "structures": ["17B20T20S20T20S20T"], "bases": [S1 + "A" * 20 + "C" * 20 + "A" * 20 + "C" * 20 + "A" * 20],
     "delimited": S1, "template_bases": 60},
]

out = {"samples": SAMPLES, "max_mismatches": 2, "min_delta": 1, "free_ns": 2, "max_no_calls": 1,
       "matchers": ["cached-hamming-distance", "pre-compute"], "expected_sample_index": 0, "qual_byte": 33,
       "cases": CASES}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_extraction_cases.json"), "w") as f:
    json.dump(out, f, indent=1)
print(f"wrote {len(CASES)} cases")
