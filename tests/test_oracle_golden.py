"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests
(tests/golden/reference_known_answers.json, transcribed from /root/reference/src/lib tests)."""
import itertools
import random

import numpy as np
import pytest

from oracle import oracle as orc


def test_hamming_distance_known_answers(golden):
    L = orc.lib()
    for a, b, free_ns, want in golden["hamming_distance"]:
        assert orc.hamming_distance(a.encode(), b.encode(), free_ns) == want
        assert L.orc_hamming_distance(a.encode(), len(a), b.encode(), len(b), free_ns) == want


def test_should_reject_delta_truth_table(golden):
    L = orc.lib()
    for delta, min_delta, want in golden["should_reject_delta"]:
        assert orc.should_reject_delta(delta, min_delta) == want
        assert bool(L.orc_should_reject_delta(delta, min_delta)) == want


def test_select_matcher_rule():
    # run.rs:196-198
    L = orc.lib()
    for n in (1, 8, 12):
        assert L.orc_select_matcher(n, -1) == orc.MATCHER_PRECOMPUTE == orc.select_matcher(n)
    for n in (13, 16, 20, 24):
        assert L.orc_select_matcher(n, -1) == orc.MATCHER_CACHED_HAMMING == orc.select_matcher(n)
        assert L.orc_select_matcher(n, 1) == orc.MATCHER_PRECOMPUTE == orc.select_matcher(n, 1)
    assert L.orc_select_matcher(8, 0) == orc.MATCHER_PRECOMPUTE  # the override cannot undo the <= 12 rule


def test_precompute_golden_map(golden):
    g = golden["precompute_map"]
    samples = [s.encode() for s in g["samples"]]
    want = {k.encode(): tuple(v) for k, v in g["expected"].items()}
    assert orc.literal_build_map(samples, g["max_mismatches"], g["min_delta"]) == want
    cfg = orc.Config(len(samples), 3, g["max_mismatches"], g["min_delta"], matcher=orc.MATCHER_PRECOMPUTE)
    co = orc.COracle(cfg, samples)
    assert co.has_map and co.map() == want
    (m1, d1), (m2, d2) = g["degenerate_equal"]
    assert orc.literal_build_map(samples, m1, d1) == orc.literal_build_map(samples, m2, d2)
    a = orc.COracle(orc.Config(2, 3, m1, d1, matcher=orc.MATCHER_PRECOMPUTE), samples).map()
    b = orc.COracle(orc.Config(2, 3, m2, d2, matcher=orc.MATCHER_PRECOMPUTE), samples).map()
    assert a == b == orc.literal_build_map(samples, m1, d1)


def _check_find(find, q):
    observed, is_match, idx, dist = q
    got = find(observed.encode())
    if is_match:
        assert got is not None, q
        if idx is not None:
            assert got[0] == idx, q
        if dist is not None:
            assert got[1] == dist, q
    else:
        assert got is None, q


def test_matcher_known_answers(golden):
    for case in golden["matcher_cases"]:
        samples = [s.encode() for s in case["samples"]]
        for kind in case["matchers"]:
            cfg = orc.Config(len(samples), len(samples[0]), case["max_mismatches"], case["min_delta"],
                             case["free_ns"], matcher=kind)
            lit = orc.LiteralDemux(cfg, samples)
            c_map = orc.COracle(cfg, samples)
            c_closed = orc.COracle(cfg, samples, force_closed_form=True)
            for q in case["queries"]:
                _check_find(lit.find, q)
                _check_find(c_map.find, q)
                _check_find(c_closed.find, q)


def test_header_barcodes_lose_their_first_delimiter(golden):
    """demux.rs:42-53 / 1779-1799: `index1+index2` in the FASTQ header -> the concatenated barcode of the case."""
    import sgdx_loader
    sg = sgdx_loader.load()
    seen = 0
    for case in golden["demux_cases"]:
        for hb, read in zip(case.get("header_barcodes", []), case["reads"]):
            assert sg.from_delimited(hb.encode()) == read.encode(), case["src"]
            seen += 1
    assert seen >= 1
    assert sg.from_delimited(b"ACGT") == b"ACGT" and sg.from_delimited(b"AC+GT+TT") == b"ACGT+TT"


def test_demux_known_answers(golden):
    for case in golden["demux_cases"]:
        samples = [s.encode() for s in case["samples"]]
        S, L = len(samples), len(samples[0])
        reads = [r.encode() for r in case["reads"]]
        for kind in case["matchers"]:
            cfg = orc.Config(S, L, case["max_mismatches"], case["min_delta"], case["free_ns"],
                             case["max_no_calls"], kind, case["index1_len"], case["index2_len"])
            lit = orc.LiteralDemux(cfg, samples).run(reads)
            arr = np.frombuffer(b"".join(reads), dtype=np.uint8).reshape(len(reads), L)
            for closed in (False, True):
                for threads, cache in ((1, False), (2, True)):
                    got = orc.COracle(cfg, samples, force_closed_form=closed).run(arr, threads, cache)
                    for res in (lit, got):
                        assert list(res["idx"]) == case["expected_idx"], case["src"]
                        if "expected_dist" in case:
                            want = [orc.NO_DIST if d is None else d for d in case["expected_dist"]]
                            assert list(res["dist"]) == want, case["src"]
                        if "per_sample" in case:
                            assert res["per_sample"].tolist() == case["per_sample"], case["src"]
                        if "per_sample_totals" in case:
                            assert res["per_sample"][:, 0].tolist() == case["per_sample_totals"], case["src"]
                        if "total_templates" in case:
                            assert res["total_templates"] == case["total_templates"]
                        if "hops" in case:
                            assert {k.decode(): v for k, v in res["hops"].items()} == case["hops"], case["src"]
                    if "unmatched" in case:
                        um = orc.unmatched_counts(arr, got["idx"], S)
                        assert {k.decode(): v for k, v in um.items()} == case["unmatched"], case["src"]
                        assert {k.decode(): v for k, v in lit["unmatched"].items()} == case["unmatched"]


def _closed_form_py(read, samples, M, D):
    i, d = orc.C.c_size_t(0), orc.C.c_size_t(0)
    ok = orc.lib().orc_find_precompute_closed_form(read, len(read), b"".join(samples), len(samples),
                                                   len(samples[0]), M, D, orc.C.byref(i), orc.C.byref(d))
    return (i.value, d.value) if ok else None


@pytest.mark.parametrize("L", [3, 4, 5])
def test_precompute_closed_form_equals_literal_map_exhaustive(L):
    """The closed form used for long barcodes must equal the literally built map on EVERY
    string over ACGTN plus a foreign byte, for random tables and all small (M, delta)."""
    rng = random.Random(1234 + L)
    alphabet = b"ACGTN"
    all_reads = [bytes(p) for p in itertools.product(alphabet, repeat=L)]
    all_reads += [b"R" + r[1:] for r in all_reads[:50]] + [r.lower() for r in all_reads[:20]]
    for trial in range(6 if L < 5 else 3):
        S = rng.randint(1, 5)
        samples = []
        while len(samples) < S:
            s = bytes(rng.choice(b"ACGT") for _ in range(L))
            if s not in samples:
                samples.append(s)
        for M, D in [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2), (2, 1), (1, 3), (2, 2), (L, 0), (L + 2, 1)]:
            lit = orc.literal_build_map(samples, M, D)
            cfg = orc.Config(S, L, M, D, matcher=orc.MATCHER_PRECOMPUTE)
            cm = orc.COracle(cfg, samples)
            assert cm.has_map and cm.map() == lit, (samples, M, D)
            for r in all_reads:
                assert _closed_form_py(r, samples, M, D) == lit.get(r), (samples, M, D, r)


def test_matchers_differ_where_the_survey_says():
    """defaults M=1, delta=2; best=1, next=3 -> PreCompute matches, Hamming rejects."""
    samples = [b"AAAAAA", b"AATTTA"]
    read = b"CAAAAA"  # d=1 to s0, d=4 to s1 -> both match
    assert orc.find_independently(read, samples, 1, 2, 0) == (0, 1)
    samples = [b"AAAAAA", b"CATTAA"]  # read CAAAAA: d=1 to s0, d=2 to s1
    assert orc.find_independently(read, samples, 1, 2, 0) is None
    samples = [b"AAAAAA", b"CTTTAA"]  # d=1, d=3: competitor exactly at M+delta
    assert orc.find_independently(read, samples, 1, 2, 0) is None
    assert orc.literal_build_map(samples, 1, 2).get(read) == (0, 1)
    assert _closed_form_py(read, samples, 1, 2) == (0, 1)


def test_cached_and_uncached_agree_and_threads_merge_exactly():
    rng = np.random.default_rng(7)
    S, L = 24, 16
    table = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(S, L))
    reads = table[rng.integers(0, S, 20000)].copy()
    mut = rng.random(reads.shape) < 0.03
    reads[mut] = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=int(mut.sum()))
    cfg = orc.Config(S, L, 1, 2, 1, -1, orc.MATCHER_CACHED_HAMMING, 8, 8)
    o = orc.COracle(cfg, table)
    base = o.run(reads, 1, False)
    for threads, cache in ((1, True), (3, False), (4, True)):
        got = o.run(reads, threads, cache)
        assert np.array_equal(got["idx"], base["idx"]) and np.array_equal(got["dist"], base["dist"])
        assert np.array_equal(got["per_sample"], base["per_sample"])
        assert got["hops"] == base["hops"] and got["total_templates"] == base["total_templates"] == 20000
    lit = orc.LiteralDemux(cfg, [bytes(t) for t in table]).run([bytes(r) for r in reads[:3000]])
    small = o.run(reads[:3000], 2, True)
    assert np.array_equal(lit["idx"], small["idx"]) and np.array_equal(lit["dist"], small["dist"])
    assert np.array_equal(lit["per_sample"], small["per_sample"]) and lit["hops"] == small["hops"]
