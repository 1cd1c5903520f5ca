"""CPU checks of the stages either side of the assignment: the oracle restatements against the reference's own
known answers, the read-structure mirror, struct layouts of the new C-ABI types."""
import ctypes as C

import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
RS = sg.ReadStructure.from_str


def _segs(rs):
    return [(g.offset, g.length, g.kind) for g in rs.segments]


def test_base_qual_oracle_reproduces_reference_test():
    """run.rs:1040-1157 (test_demux_filtering...): read structure 17B39T, qualities '?'*17 + bytes 35..73.
    Expected per read: 39 template bases, 11 with q >= 30, 21 with q >= 20; 4 / 2 / 1 reads kept."""
    one = [63] * 17 + list(range(35, 74))
    for reads, (bases, q30, q20) in ((4, (156, 44, 84)), (2, (78, 22, 42)), (1, (39, 11, 21))):
        q = np.array([one] * reads, dtype=np.uint8)
        out, short = orc.base_qual_counts(q, None, _segs(RS("17B39T")), np.zeros(reads, dtype=np.uint16), 4)
        assert (int(out[0, 4]), int(out[0, 1]), int(out[0, 0])) == (bases, q30, q20) and short == 0
        assert (int(out[0, 2]), int(out[0, 3])) == (30 * 17 * reads, 17 * reads)  # metrics.rs:500-505
        assert int(out[1:].sum()) == 0


def test_base_qual_oracle_segment_kinds_and_wrap():
    """metrics.rs:499-518: only T adds to bases/q20/q30, only B to the index sums; M and S add nothing.
    `q - 33` on a u8 wraps for bytes below '!' (release build): 32 -> 255 counts as >= Q30."""
    q = np.array([[33 + 19, 33 + 20, 33 + 29, 33 + 30, 32, 33 + 40, 33 + 2, 33 + 3]], dtype=np.uint8)
    out, _ = orc.base_qual_counts(q, None, _segs(RS("5T1M1S1B")), np.array([1], dtype=np.uint16), 2)
    assert out[1].tolist() == [4, 2, 3, 1, 5]
    out, short = orc.base_qual_counts(q, np.array([6], dtype=np.uint32), _segs(RS("5T1M+T")), np.array([0], dtype=np.uint16), 2)
    assert out[0].tolist() == [4, 2, 0, 0, 5] and short == 0        # `+T` is empty at length 6
    out, short = orc.base_qual_counts(q, np.array([5], dtype=np.uint32), _segs(RS("5T1M+T")), np.array([0], dtype=np.uint16), 2)
    assert int(out.sum()) == 0 and short == 1                        # shorter than the fixed segments


def test_unmatched_literal_reproduces_reference_tests(golden):
    """run.rs:652-689: three unmatched barcodes, each once; run.rs:1444-1451: the hopped barcode once."""
    for case in golden["demux_cases"]:
        if "unmatched" not in case:
            continue
        table = [s.encode() for s in case["samples"]]
        L = len(table[0])
        il = (case["index1_len"], case["index2_len"]) if case["index1_len"] else (0, 0)
        for k in case["matchers"]:
            cfg = orc.Config(len(table), L, case["max_mismatches"], case["min_delta"], case["free_ns"],
                             -1 if case["max_no_calls"] is None else case["max_no_calls"],
                             orc.MATCHER_PRECOMPUTE if k == 1 else orc.MATCHER_CACHED_HAMMING, il[0], il[1])
            res = orc.LiteralDemux(cfg, table).run([r.encode() for r in case["reads"]])
            c = orc.UnmatchedCounterLiteral()
            for r, i in zip(case["reads"], res["idx"]):
                if i == len(table):
                    c.insert(r.encode())
            assert {k_.decode(): v for k_, v in c.unmatched_counter.items()} == case["unmatched"]
            assert len(c.top(1000)) == len(case["unmatched"])


def test_unmatched_literal_downsizes_like_the_reference():
    """metrics.rs:151-175: the map is cut when it already holds max_counter_size keys and one more arrives."""
    c = orc.UnmatchedCounterLiteral(max_counter_size=4, downsize_to=2)
    for k in [b"A", b"A", b"A", b"C", b"C", b"G", b"T"]:
        c.insert(k)
    assert len(c.unmatched_counter) == 4
    c.insert(b"A")                      # present already, but the map is full: cut first (reference behaviour)
    assert c.unmatched_counter == {b"A": 4, b"C": 2}
    assert c.top(1) == [(b"A", 4)]


def test_extract_and_route_oracles():
    i1 = np.frombuffer(b"AAAACCCC" b"GGGGTTTT", dtype=np.uint8).reshape(2, 8)
    r1 = np.frombuffer(b"NNACGTxx" b"NNTGCAyy", dtype=np.uint8).reshape(2, 8)
    got = orc.extract_sample_barcodes([r1, i1], [_segs(RS("2S4B+T")), _segs(RS("8B"))])
    assert [bytes(r) for r in got] == [b"ACGTAAAACCCC", b"TGCAGGGGTTTT"]  # FASTQ order, then segment order
    idx = np.array([2, 0, 2, 1, 0, 2], dtype=np.uint16)
    order, off = orc.route_reads(idx, 2)
    assert order.tolist() == [1, 4, 3, 0, 2, 5] and off.tolist() == [0, 2, 3, 6]


def test_read_structure_mirror():
    rs = RS("8B+T")
    assert [(g.offset, g.length, g.kind) for g in rs.segments] == [(0, 8, "B"), (8, None, "T")]
    assert rs.sample_barcode_length() == 8 and rs.min_length() == 8
    assert RS("6M+T").sample_barcode_length() == 0
    assert RS("3b2s4B").sample_barcode_length() == 7
    for bad in ("", "8", "B8", "+T8B", "0T", "8Q", "8B +T x"):
        with pytest.raises(ValueError):
            RS(bad)
    arr, n = sg.read_structure.c_segments([RS("10B"), RS("4S+T")])
    assert n == 3 and (arr[2].source, arr[2].offset, arr[2].length, arr[2].kind) == (1, 4, 0xFFFFFFFF, ord("T"))


def test_new_struct_layouts_match_header():
    from singular_demux_b200 import _lib
    assert C.sizeof(_lib.UnmatchedInfo) == 4 * 8
    assert C.sizeof(_lib.Segment) == 4 * 4
    assert C.sizeof(_lib.BaseQual) == 5 * 8


def test_post_calls_validate_before_touching_cuda():
    L = sg.lib()
    assert L.sgdx_unmatched_enable(None, 0, 0) == -1
    assert L.sgdx_route_device(None, 0, None, 0, None, None) == -1
    assert L.sgdx_base_qual_counters(None, None, None) == -1
    assert L.sgdx_qual_counts_device(None, 0, None, 0, 150, None, None, 0, None) == -1


def test_literal_demultiplexer_reproduces_reference_run():
    """run.rs:1040-1157 through the read-by-read restatement of demux.rs:445-584: four 17B39T reads of sample 1
    -> 4 templates, 156 bases, 44 >= Q30, 84 >= Q20 on sample 1, nothing anywhere else."""
    table = [b"AAAAAAAAGATTACAGA", b"CCCCCCCCGATTACAGA", b"GGGGGGGGGATTACAGA", b"GGGGGGTTGATTACAGA"]
    cfg = orc.Config(4, 17, 2, 1, 1, 1, orc.MATCHER_CACHED_HAMMING, 0, 0)
    bases = np.frombuffer((table[0] + b"A" * 39) * 4, dtype=np.uint8).reshape(4, 56)
    quals = np.array([[63] * 17 + list(range(35, 74))] * 4, dtype=np.uint8)
    res = orc.LiteralDemultiplexer(cfg, table).run([(bases, quals, None)], [_segs(RS("17B39T"))])
    assert res["idx"].tolist() == [0, 0, 0, 0] and res["total_templates"] == 4
    assert res["per_sample"][0].tolist() == [4, 4, 0]
    assert res["base_qual"][0].tolist() == [84, 44, 17 * 30 * 4, 17 * 4, 156] and int(res["base_qual"][1:].sum()) == 0
    assert res["per_sample_reads"][0] == [0, 1, 2, 3] and res["unmatched"] == {}


def test_vectorised_demultiplex_equals_the_literal_one():
    """The composed numpy/C oracle against the read-by-read one on a small ragged dual-index paired-end batch."""
    rng = np.random.default_rng(12)
    w = sg.synth.Workload("lit", 12, 12, 6, 1, 1, n_reads=600, table_seed=5, read_seed=6, p_true=0.7, p_hop=0.1, p_n=0.02)
    table = w.table()
    bc = w.reads_host(0, 600, table, threads=1)
    n = 600
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    r1 = letters[rng.integers(0, 5, (n, 30))].copy()
    r2 = letters[rng.integers(0, 5, (n, 25))].copy()
    i1, i2 = bc[:, :6].copy(), letters[rng.integers(0, 5, (n, 9))].copy()
    i2[:, 3:9] = bc[:, 6:]
    fq = []
    for a in (i1, i2, r1, r2):
        q = rng.integers(33, 75, a.shape, dtype=np.uint8)
        q[rng.random(a.shape) < 0.01] = 20  # below '!': wraps
        fq.append([a, q, None])
    fq[2][2] = rng.integers(4, 31, n).astype(np.uint32)   # R1 ragged: 4M+T needs >= 4
    structures = [_segs(RS(s)) for s in ("6B", "3S6B", "4M+T", "25T")]
    cfg = orc.Config(12, 12, 1, 1, 1, -1, orc.MATCHER_CACHED_HAMMING, 6, 6)
    lit = orc.LiteralDemultiplexer(cfg, [bytes(t) for t in table]).run([tuple(f) for f in fq], structures)
    vec = orc.demultiplex_batch(cfg, [bytes(t) for t in table], [tuple(f) for f in fq], structures)
    assert np.array_equal(lit["idx"], vec["idx"]) and np.array_equal(lit["dist"], vec["dist"])
    assert np.array_equal(lit["per_sample"], vec["per_sample"]) and lit["hops"] == vec["hops"]
    assert lit["unmatched"] == vec["unmatched"] and len(lit["unmatched"]) > 10
    assert np.array_equal(lit["base_qual"], vec["base_qual"])
    for s in range(13):
        assert lit["per_sample_reads"][s] == vec["order"][int(vec["offsets"][s]):int(vec["offsets"][s + 1])].tolist()
    assert [bytes(b) for b in vec["barcodes"]] == lit["barcodes"]


def _extraction_golden():
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "reference_extraction_cases.json")) as f:
        return json.load(f)


def _extraction_case_inputs(g, case):
    rss = [RS(s) for s in case["structures"]]
    fastqs = []
    for b in case["bases"]:
        a = np.frombuffer(b.encode(), dtype=np.uint8).reshape(1, -1)
        fastqs.append((a, np.full(a.shape, g["qual_byte"], dtype=np.uint8), None))
    return rss, fastqs


def test_oracle_reproduces_the_reference_extraction_tests():
    """demux.rs:867-1043: five read-structure layouts, every one lands in sample 1; the concatenated barcode is
    SAMPLE_BARCODE_1, its delimited form the header value the reference asserts."""
    g = _extraction_golden()
    table = [s.encode() for s in g["samples"]]
    for case in g["cases"]:
        rss, fastqs = _extraction_case_inputs(g, case)
        bc = orc.extract_sample_barcodes([f[0] for f in fastqs], [_segs(rs) for rs in rss])
        assert bytes(bc[0]) == table[0], case["src"]
        cuts, pos = [], 0  # delimited form: B segments joined by '+' (demux.rs:504-522)
        for rs, (a, _, _) in zip(rss, fastqs):
            for s in rs.segments:
                if s.kind == "B":
                    ln = s.length if s.length is not None else a.shape[1] - s.offset
                    cuts.append(bytes(bc[0][pos:pos + ln]).decode())
                    pos += ln
        assert "+".join(cuts) == case["delimited"], case["src"]
        for k in (orc.MATCHER_CACHED_HAMMING, orc.MATCHER_PRECOMPUTE):
            if k == orc.MATCHER_PRECOMPUTE:
                continue  # the literal PreCompute map of a 17-mer table is too large to build in a unit test
            cfg = orc.Config(4, 17, g["max_mismatches"], g["min_delta"], g["free_ns"], g["max_no_calls"], k, 0, 0)
            res = orc.LiteralDemultiplexer(cfg, table).run(fastqs, [_segs(rs) for rs in rss])
            assert res["idx"].tolist() == [g["expected_sample_index"]], case["src"]
            assert res["per_sample_reads"][0] == [0] and all(not r for r in res["per_sample_reads"][1:])
            assert res["base_qual"][0].tolist() == [0, 0, 0, 17, case["template_bases"]], case["src"]


def test_delimited_unmatched_keys_have_a_plus_between_every_barcode_segment():
    """ExtractedBarcode::add_bases (demux.rs:56-62) pushes `+` before every B segment but the first -- several B
    segments inside one read (4B2S4B) and three barcode reads (10B+T 8B 8B) included -- and that delimited string is
    what the unmatched counter receives (demux.rs:553-555) and most_frequent_unmatched.tsv shows."""
    rs = [sg.ReadStructure.from_str(s) for s in ("4B2S4B", "+T")]
    assert sg.barcode_segment_lengths(rs, read_lens=[10, 50]) == [4, 4]
    assert sg.delimit_barcode(b"ACGTTTGA", [4, 4]) == b"ACGT+TTGA"
    rs3 = [sg.ReadStructure.from_str(s) for s in ("10B+T", "8B", "8B")]
    cuts = sg.barcode_segment_lengths(rs3, read_lens=[160, 8, 8])
    assert cuts == [10, 8, 8]
    key = b"A" * 10 + b"C" * 8 + b"G" * 8
    assert sg.delimit_barcode(key, cuts) == b"AAAAAAAAAA+CCCCCCCC+GGGGGGGG"
    assert sg.delimit_barcode(b"ACGTACGT", [8]) == b"ACGTACGT" == sg.delimit_barcode(b"ACGTACGT", None)
    assert sg.barcode_segment_lengths([sg.ReadStructure.from_str("6B2M+T"), sg.ReadStructure.from_str("+B")], [100, 9]) == [6, 9]
    # the literal walk of demultiplex produces exactly these pieces
    fq = [np.frombuffer(b"ACGTNNTTGA", dtype=np.uint8).reshape(1, -1)]
    bc = orc.extract_sample_barcodes(fq, [[(0, 4, "B"), (4, 2, "S"), (6, 4, "B")]])
    assert sg.delimit_barcode(bytes(bc[0]), [4, 4]) == b"ACGT+TTGA"
    # the counters' rows: three segments, and the two-read shorthand
    c = sg.UnmatchedCounter()
    c.insert_counts(np.frombuffer(key, dtype=np.uint8).reshape(1, -1), np.array([3]))
    assert c.most_frequent(5, segment_lengths=cuts) == [sg.BarcodeCount("AAAAAAAAAA+CCCCCCCC+GGGGGGGG", 3)]
    assert c.most_frequent(5, index1_length=10) == [sg.BarcodeCount("AAAAAAAAAA+CCCCCCCCGGGGGGGG", 3)]
    assert c.most_frequent(5) == [sg.BarcodeCount(key.decode(), 3)]
    rows = sg.sharding.merge_unmatched([{key: 2}, {key: 5}], n=3, segment_lengths=cuts)
    assert rows == [sg.BarcodeCount("AAAAAAAAAA+CCCCCCCC+GGGGGGGG", 7)]
