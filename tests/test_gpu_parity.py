"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.
Run on a B200: python -m pytest tests -m gpu."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
pytestmark = pytest.mark.gpu

KERNELS = [sg.KERNEL_DIRECT, sg.KERNEL_SEEDED]
ORC_KIND = {sg.MatcherKind.CachedHammingDistance: orc.MATCHER_CACHED_HAMMING, sg.MatcherKind.PreCompute: orc.MATCHER_PRECOMPUTE}
CLS = {sg.MatcherKind.CachedHammingDistance: sg.CachedHammingDistanceMatcher, sg.MatcherKind.PreCompute: sg.PreComputeMatcher}


def _samples(table):
    out = [sg.SampleMetadata(f"s{i}", bytes(b), i) for i, b in enumerate(table)]
    out.append(sg.SampleMetadata(sg.UNDETERMINED_NAME, b"N" * len(out[0].barcode), len(out)))
    return out


def _stage(table, kind, M, D, F, mnc, il, kernel, slots=2):
    return sg.BarcodeAssignmentStage(_samples(table), CLS[kind], M, D, F, mnc, il, kernel=kernel, slots=slots)


def _oracle(table, kind, M, D, F, mnc, il, reads, **kw):
    S, L = len(table), len(table[0])
    cfg = orc.Config(S, L, M, D, F, -1 if mnc is None else mnc, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
    return orc.COracle(cfg, [bytes(t) for t in table], **kw).run(reads, threads=8, use_cache=True)


def _assert_same(stage, got, want, n):
    assert np.array_equal(got.sample_indices, want["idx"])
    assert np.array_equal(got.hamming_dists, want["dist"])
    m = stage.metrics()
    assert np.array_equal(m.per_sample_array(), want["per_sample"])
    assert m.total_templates == want["total_templates"] == n
    if stage.index_lengths:
        assert m.sample_barcode_hop_tracker.count_tracker == want["hops"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_reference_known_answers_on_device(golden, kernel):
    """Every demux-level known answer of the reference's own tests, through the CUDA path."""
    for case in golden["demux_cases"]:
        table = [s.encode() for s in case["samples"]]
        S, L = len(table), len(table[0])
        reads = np.frombuffer("".join(case["reads"]).encode(), dtype=np.uint8).reshape(-1, L)
        il = (case["index1_len"], case["index2_len"]) if case["index1_len"] else None
        for k in case["matchers"]:
            kind = sg.MatcherKind.PreCompute if k == 1 else sg.MatcherKind.CachedHammingDistance
            st = _stage(table, kind, case["max_mismatches"], case["min_delta"], case["free_ns"], case["max_no_calls"], il, kernel)
            got = st.demultiplex(reads)
            assert got.sample_indices.tolist() == case["expected_idx"], case["src"]
            if "expected_dist" in case:
                assert got.hamming_dists.tolist() == [0xFF if d is None else d for d in case["expected_dist"]]
            m = st.metrics()
            if "per_sample" in case:
                assert m.per_sample_array().tolist() == case["per_sample"]
            if "per_sample_totals" in case:
                assert m.per_sample_array()[:, 0].tolist() == case["per_sample_totals"]
            if "total_templates" in case:
                assert m.total_templates == case["total_templates"]
            if "hops" in case:
                assert {k_.decode(): v for k_, v in m.sample_barcode_hop_tracker.count_tracker.items()} == case["hops"]
                if case["hops"]:
                    assert m.sample_barcode_hop_tracker.barcode_counts()[0] == sg.BarcodeCount("TTT+AAA", 1)
            if "unmatched" in case:
                c = sg.demux.count_unmatched(got.unmatched)
                assert {k_.decode(): v for k_, v in c.counts.items()} == case["unmatched"]
            st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_matcher_find_known_answers_on_device(golden, kernel):
    for case in golden["matcher_cases"]:
        table = [s.encode() for s in case["samples"]]
        for k in case["matchers"]:
            kind = sg.MatcherKind.PreCompute if k == 1 else sg.MatcherKind.CachedHammingDistance
            m = CLS[kind](table, case["max_mismatches"], case["min_delta"], case["free_ns"], kernel=kernel)
            for observed, is_match, idx, dist in case["queries"]:
                r = m.find(observed)
                assert r.is_match() == is_match, (case["src"], observed)
                if is_match:
                    assert (r.sample_index, r.hamming_dist) == (idx, dist)
                assert r.barcode == observed.encode()
            m.close()


def _random_case(rng, S, L, i1, p_n=0.02, p_x=0.004, n=40000, near=0.0):
    w = sg.synth.Workload("r", S, L, i1, 1, 2, n_reads=n, table_seed=int(rng.integers(1 << 30)),
                          read_seed=int(rng.integers(1 << 30)), p_true=0.80, p_hop=0.08 if i1 else 0.0,
                          p_sub=0.03, p_n=p_n, p_x=p_x, min_dist=1 if S > 4 ** min(L, 4) // 8 else 2)
    table = w.table()
    reads = w.reads_host(0, n, table)
    if i1:  # exercise the reversed-orientation hop rule and Undetermined's all-N halves
        k = n // 50
        a, b = rng.integers(0, S, k), rng.integers(0, S, k)
        if i1 * 2 == L:
            reads[:k, :i1], reads[:k, i1:] = table[b, i1:], table[a, :i1]
        reads[k:k + 20, :i1] = ord("N")
        reads[k + 20:k + 40, i1:] = ord("N")
        reads[k + 40:k + 50, :] = ord("N")
    return w, table, reads


PARAMS = [  # (M, delta, free_ns, max_no_calls)
    (1, 2, 1, None), (0, 0, 0, None), (2, 1, 2, 1), (3, 1, 1, 0), (1, 0, 0, 2), (2, 3, 1, None), (40, 0, 3, None),
    (1, 70, 1, None),
]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("S,L,i1", [(1, 8, 0), (2, 4, 2), (7, 12, 0), (24, 16, 8), (96, 16, 8), (97, 20, 10), (384, 20, 10),
                                    (300, 24, 12), (50, 32, 16), (33, 13, 5), (64, 6, 3), (3, 1, 0), (130, 31, 0)])
def test_random_tables_and_reads_match_oracle(kernel, S, L, i1):
    rng = np.random.default_rng(S * 1000 + L)
    w, table, reads = _random_case(rng, S, L, i1)
    il = w.index_lengths
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for (M, D, F, mnc) in PARAMS:
            st = _stage(table, kind, M, D, F, mnc, il, kernel)
            got = st.demultiplex(reads)
            want = _oracle(table, kind, M, D, F, mnc, il, reads, force_closed_form=True)
            _assert_same(st, got, want, reads.shape[0])
            st.close()


def test_seeded_kernel_engages_and_falls_back():
    """AUTO picks the seeded kernel where its index pays, SEEDED is honoured whenever an index exists, and
    configurations without one (T+1 > L) run the direct kernel."""
    t96 = sg.synth.C1.table()
    cases = [  # (table, M, D, requested, effective)
        (t96, 1, 2, sg.KERNEL_AUTO, sg.KERNEL_SEEDED), (t96, 1, 2, sg.KERNEL_DIRECT, sg.KERNEL_DIRECT),
        (t96, 1, 2, sg.KERNEL_SEEDED, sg.KERNEL_SEEDED), (t96, 9, 9, sg.KERNEL_SEEDED, sg.KERNEL_DIRECT),
        (t96, 4, 4, sg.KERNEL_SEEDED, sg.KERNEL_SEEDED), (t96, 4, 4, sg.KERNEL_AUTO, sg.KERNEL_DIRECT),
        (t96[:2], 1, 2, sg.KERNEL_AUTO, sg.KERNEL_DIRECT),
    ]
    for table, M, D, req, eff in cases:
        m = sg.CachedHammingDistanceMatcher([bytes(t) for t in table], M, D, 1, kernel=req)
        assert m.kernel == eff, (len(table), M, D, req, m.kernel)
        m.close()


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])
def test_baseline_tables_take_every_fast_path(cfg):
    """The BASELINE tables (unique dual indexes, min pairwise distance 6) must run with the seed index, the
    exact-match front end, the neighbour table and the direct hop tables all enabled; a silently disabled
    table would only show up as a slower kernel."""
    from singular_demux_b200 import _lib
    w = sg.synth.CONFIGS[cfg]
    kind = sg.select_matcher(w.barcode_len)
    m = CLS[kind]([bytes(t) for t in w.table()], w.max_mismatches, w.min_delta, w.free_ns, index_lengths=w.index_lengths)
    want = _lib.PATH_SEED_INDEX | _lib.PATH_EXACT
    if cfg == "C3":  # the neighbour stage is only enabled where the seed search it saves is long
        want |= _lib.PATH_NEIGHBOURS
    if max(w.index_lengths) <= 11:
        want |= _lib.PATH_HOP_DIRECT
    want |= _lib.PATH_PERFECT_HASH  # compact records (C3: 16-byte records): the perfect hash of the match neighbourhood
    if cfg == "C3":                 # ... which stops at the barcodes themselves there: second level + short seed index
        want |= _lib.PATH_SHORT_SEED
    assert m.path_flags() == want and m.kernel_name() == "sgdx_match_exact_first"
    m.close()


@pytest.mark.parametrize("p_n,p_x", [(0.03, 0.0), (0.10, 0.01), (0.30, 0.05)])
@pytest.mark.parametrize("S,L,i1,M,D,F", [(384, 20, 10, 1, 2, 1), (200, 24, 12, 3, 1, 2), (96, 16, 8, 2, 1, 3),
                                            (1536, 24, 12, 3, 1, 1), (60, 32, 16, 2, 2, 4), (500, 12, 6, 1, 1, 1)])
def test_seeded_no_call_paths(S, L, i1, M, D, F, p_n, p_x):
    """No-call heavy reads: windows with one invalid base (4 buckets), windows with several (warp-cooperative
    scan of the queue) and reads that cannot match at all, for both rule sets and with the N gate."""
    rng = np.random.default_rng(S + L + int(p_n * 100))
    w, table, reads = _random_case(rng, S, L, i1, p_n=p_n, p_x=p_x, n=60000)
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for mnc in (None, 2):
            st = _stage(table, kind, M, D, F, mnc, w.index_lengths, sg.KERNEL_SEEDED)
            assert st.matcher.kernel == sg.KERNEL_SEEDED
            got = st.demultiplex(reads)
            want = _oracle(table, kind, M, D, F, mnc, w.index_lengths, reads, force_closed_form=True)
            _assert_same(st, got, want, reads.shape[0])
            st.close()


@pytest.mark.parametrize("neighbours", [False, True])
@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_exact_and_neighbour_front_end_rule_grid(cfg, neighbours, monkeypatch):
    """Well separated tables (min pairwise distance 6) take the exact-match and (forced here, it is a cost
    heuristic otherwise) one-substitution shortcuts of the seeded kernel; every rule parameter that touches
    those shortcuts is swept against the oracle."""
    from singular_demux_b200 import _lib
    if neighbours:
        monkeypatch.setenv("SGDX_FORCE_NEIGHBOURS", "1")
    base = sg.synth.CONFIGS[cfg]
    w = sg.synth.Workload("grid", base.num_samples, base.barcode_len, base.index1_len, 1, 2, n_reads=40_000,
                          table_seed=base.table_seed, read_seed=77, p_true=0.85, p_hop=0.05, p_sub=0.03, p_n=0.02,
                          p_x=0.002)
    table = w.table()
    reads = w.reads_host(0, w.n_reads, table)
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for M in (0, 1, 2):
            for D in (0, 1, 2, 3, 4):
                for F, mnc in ((0, None), (1, None), (2, 0), (1, 1), (0, 0)):
                    st = _stage(table, kind, M, D, F, mnc, w.index_lengths, sg.KERNEL_SEEDED)
                    if neighbours and D <= 3 and st.matcher.path_flags() & _lib.PATH_EXACT:  # dmin = 6 > D + 2
                        assert st.matcher.path_flags() & _lib.PATH_NEIGHBOURS
                    got = st.demultiplex(reads)
                    want = _oracle(table, kind, M, D, F, mnc, w.index_lengths, reads, force_closed_form=True)
                    _assert_same(st, got, want, reads.shape[0])
                    st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_precompute_against_literal_map_oracle(kernel):
    """Short barcodes: the oracle side uses the literally constructed PreCompute map, not the closed form."""
    rng = np.random.default_rng(99)
    for (S, L, i1) in [(3, 5, 0), (12, 8, 4), (1, 8, 0), (40, 10, 5)]:
        w, table, reads = _random_case(rng, S, L, i1, p_n=0.05, p_x=0.01, n=30000)
        for (M, D) in [(1, 2), (1, 1), (2, 1), (0, 0), (1, 3), (2, 2)]:
            st = _stage(table, sg.MatcherKind.PreCompute, M, D, 1, None, w.index_lengths, kernel)
            got = st.demultiplex(reads)
            S_, L_ = len(table), len(table[0])
            cfg = orc.Config(S_, L_, M, D, 1, -1, orc.MATCHER_PRECOMPUTE, i1, L - i1 if i1 else 0)
            o = orc.COracle(cfg, [bytes(t) for t in table])
            assert o.has_map or M + D == 0  # M = delta = 0 makes pass 2 enumerate all 5^L strings
            _assert_same(st, got, o.run(reads, threads=8), reads.shape[0])
            st.close()


def test_synth_device_equals_host():
    import torch
    w = sg.synth.Workload("t", 96, 20, 10, 1, 2, p_n=0.01, p_x=0.002)
    table = w.table()
    n, first = 100_003, 777
    d_table = torch.from_numpy(table).cuda()
    d_out = torch.empty((n, 20), dtype=torch.uint8, device="cuda")
    w.reads_device(first, n, d_table.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), w.reads_host(first, n, table))


@pytest.mark.parametrize("kernel", KERNELS)
def test_device_resident_and_packed_paths_equal_host_path(kernel):
    import torch
    rng = np.random.default_rng(5)
    for (S, L, i1) in [(96, 16, 8), (200, 20, 10), (31, 9, 0)]:
        w, table, reads = _random_case(rng, S, L, i1, n=50001)
        n = reads.shape[0]
        st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, kernel)
        host = st.demultiplex(reads)
        base = st.metrics().per_sample_array()
        m = st.matcher
        d_in = torch.from_numpy(reads).cuda()
        d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
        d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
        m.set_stream(0, torch.cuda.current_stream().cuda_stream)
        m.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_idx.cpu().numpy().view(np.uint16), host.sample_indices)
        assert np.array_equal(d_dist.cpu().numpy(), host.hamming_dists)
        d_packed = torch.empty((n, 4), dtype=torch.int32, device="cuda")
        d_idx.zero_()
        d_dist.zero_()
        m.pack_device(d_in.data_ptr(), n, d_packed.data_ptr())
        m.match_packed_device(d_packed.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_idx.cpu().numpy().view(np.uint16), host.sample_indices)
        assert np.array_equal(d_dist.cpu().numpy(), host.hamming_dists)
        assert np.array_equal(st.metrics().per_sample_array(), base * 3)  # counters are cumulative
        # an unaligned device pointer (odd byte offset) must still work
        d_shift = torch.empty(n * L + 1, dtype=torch.uint8, device="cuda")
        d_shift[1:] = d_in.view(-1)
        m.match_device(d_shift.data_ptr() + 1, n, d_idx.data_ptr(), d_dist.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_idx.cpu().numpy().view(np.uint16), host.sample_indices)
        m.set_stream(0, 0)
        st.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_slots_shards_and_merge_equal_single_pass(kernel):
    """Batches on different slots / different handles (the multi-GPU sharding model) merge to exactly the
    single-pass counters (metrics.rs:315-328)."""
    rng = np.random.default_rng(11)
    w, table, reads = _random_case(rng, 96, 16, 8, n=120_000)
    n = reads.shape[0]
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, reads)
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, kernel, slots=3)
    idx = np.empty(n, dtype=np.uint16)
    dist = np.empty(n, dtype=np.uint8)
    cuts = [0, 1, 40_000, 40_000, 99_999, n]
    for i in range(len(cuts) - 1):
        a, b = cuts[i], cuts[i + 1]
        st.matcher.submit(reads[a:b], idx[a:b], dist[a:b], slot=i % 3)
    for s in range(3):
        st.matcher.sync(s)
    assert np.array_equal(idx, want["idx"]) and np.array_equal(dist, want["dist"])
    assert np.array_equal(st.metrics().per_sample_array(), want["per_sample"])
    st.close()
    merged = None
    for part in np.array_split(np.arange(n), 4):
        s2 = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, kernel)
        s2.demultiplex(reads[part])
        m = s2.metrics()
        if merged is None:
            merged = m
        else:
            merged.update_with(m)
        s2.close()
    assert np.array_equal(merged.per_sample_array(), want["per_sample"])
    assert merged.total_templates == n and merged.sample_barcode_hop_tracker.count_tracker == want["hops"]


def test_back_to_back_create_and_first_batch():
    """A batch queued right after sgdx_create must see fully uploaded tables and zeroed counters (the slots'
    streams do not order against the default stream the uploads ran on): many short create -> run cycles."""
    rng = np.random.default_rng(7)
    w, table, reads = _random_case(rng, 96, 16, 8, n=40_000)
    for kind, M, D, F, mnc in [(sg.MatcherKind.PreCompute, 1, 2, 1, None), (sg.MatcherKind.PreCompute, 2, 1, 2, 1),
                               (sg.MatcherKind.CachedHammingDistance, 1, 0, 0, 2)]:
        want = _oracle(table, kind, M, D, F, mnc, w.index_lengths, reads, force_closed_form=True)
        for _ in range(12):
            st = _stage(table, kind, M, D, F, mnc, w.index_lengths, sg.KERNEL_SEEDED)
            _assert_same(st, st.demultiplex(reads), want, reads.shape[0])
            st.close()


@pytest.mark.parametrize("S,L,i1,expect", [(3000, 30, 15, "sgdx_match_seeded"), (8192, 20, 10, "sgdx_match_direct"),
                                            (1536, 24, 12, "sgdx_match_exact_first")])
def test_large_tables_pick_a_kernel_that_fits(S, L, i1, expect):
    """Table sizes around the shared-memory limits: 1536 x 24 runs the exact-first kernel with one 1024-thread
    CTA per SM, 3000 x 30 only fits the plain seeded kernel, 8192 x 16 only the direct kernel."""
    w = sg.synth.Workload("big", S, L, i1, 1, 1, n_reads=30_000, table_seed=S, read_seed=L, p_true=0.85, p_hop=0.05,
                          p_sub=0.02, p_n=0.01, p_x=0.001, min_dist=2 if S < 8192 else 1)
    table = w.table()
    reads = w.reads_host(0, w.n_reads, table)
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 1, 1, None, w.index_lengths, sg.KERNEL_AUTO)
    assert st.matcher.kernel_name() == expect
    got = st.demultiplex(reads)
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 1, 1, None, w.index_lengths, reads)
    _assert_same(st, got, want, reads.shape[0])
    st.close()


def test_two_host_threads_on_two_slots():
    """One handle, two host threads, each on its own slot (the `Matcher + Send + Sync` contract, demux.rs:422):
    results and the shared device counters must equal a single-threaded pass."""
    import threading
    rng = np.random.default_rng(21)
    w, table, reads = _random_case(rng, 96, 16, 8, n=200_000)
    n = reads.shape[0]
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 1, 1, None, w.index_lengths, reads)
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 1, 1, None, w.index_lengths, sg.KERNEL_AUTO, slots=2)
    idx = np.empty(n, dtype=np.uint16)
    dist = np.empty(n, dtype=np.uint8)
    errors = []

    def work(slot):
        try:
            lo, hi = slot * (n // 2), n if slot else n // 2
            for a in range(lo, hi, 7_001):  # many small batches so the two threads really interleave
                b = min(a + 7_001, hi)
                st.matcher.submit(reads[a:b], idx[a:b], dist[a:b], slot)
                st.matcher.sync(slot)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    ts = [threading.Thread(target=work, args=(s_,)) for s_ in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors
    assert np.array_equal(idx, want["idx"]) and np.array_equal(dist, want["dist"])
    m = st.metrics()
    assert np.array_equal(m.per_sample_array(), want["per_sample"]) and m.total_templates == n
    assert m.sample_barcode_hop_tracker.count_tracker == want["hops"]
    st.close()


def test_two_devices_merge_to_single_pass():
    """Sharding across GPUs: one handle per device, counters summed on the host (needs >= 2 GPUs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(22)
    w, table, reads = _random_case(rng, 96, 16, 8, n=100_000)
    n = reads.shape[0]
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, reads)
    parts, idx = [], np.empty(n, dtype=np.uint16)
    for dev in (0, 1):
        a, b = sg.sharding.shard_range(n, dev, 2)
        st = sg.BarcodeAssignmentStage(_samples(table), sg.CachedHammingDistanceMatcher, 1, 2, 1, None, w.index_lengths,
                                       device=dev)
        idx[a:b] = st.demultiplex(reads[a:b]).sample_indices
        parts.append(st.metrics())
        st.close()
    merged = sg.sharding.merge_metrics(parts)
    assert np.array_equal(idx, want["idx"])
    assert np.array_equal(merged.per_sample_array(), want["per_sample"]) and merged.total_templates == n
    assert merged.sample_barcode_hop_tracker.count_tracker == want["hops"]


def test_cpp_host_mirror_known_answers():
    """The reference's known-answer tests written against the C++ host mirror (tests/cpp/test_host_mirror.cpp,
    singular-demux_b200/host/sgdx_host.hpp), run as a compiled program against libsgdx.so."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_host_mirror")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "tests", "cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference known answers reproduced" in r.stdout


def test_batcher_returns_batches_in_order():
    from singular_demux_b200._lib import check, lib
    rng = np.random.default_rng(3)
    w, table, reads = _random_case(rng, 96, 16, 8, n=100_000)
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, reads)
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, sg.KERNEL_AUTO, slots=2)
    b = C.c_void_p()
    check(lib().sgdx_batcher_create(st.matcher.handle, 16_384, C.byref(b)))
    got_idx, got_dist = [], []

    def drain(everything):
        while True:
            pi, pd, n = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
            if not everything and len(got_idx) * 16_384 + 3 * 16_384 > pushed[0]:
                return
            check(lib().sgdx_batcher_next(b, C.byref(pi), C.byref(pd), C.byref(n)))
            if n.value == 0:
                return
            got_idx.append(np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint16)), (n.value,)).copy())
            got_dist.append(np.ctypeslib.as_array(C.cast(pd, C.POINTER(C.c_uint8)), (n.value,)).copy())

    pushed = [0]
    for chunk in np.array_split(reads, 100):  # 1000-read chunks like --chunksize
        c = np.ascontiguousarray(chunk)
        check(lib().sgdx_batcher_push(b, c.ctypes.data, c.shape[0]))
        pushed[0] += c.shape[0]
        drain(False)
    check(lib().sgdx_batcher_flush(b))
    drain(True)
    lib().sgdx_batcher_destroy(b)
    assert np.array_equal(np.concatenate(got_idx), want["idx"])
    assert np.array_equal(np.concatenate(got_dist), want["dist"])
    assert np.array_equal(st.metrics().per_sample_array(), want["per_sample"])
    st.close()


def test_empty_and_tiny_batches():
    st = _stage([b"ACGTACGTACGTACGT", b"TTTTACGTACGTAAAA"], sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, (8, 8), sg.KERNEL_AUTO)
    got = st.demultiplex(np.zeros((0, 16), dtype=np.uint8))
    assert got.sample_indices.shape == (0,)
    assert st.metrics().total_templates == 0
    got = st.demultiplex([b"ACGTACGTACGTACGT"])
    assert got.sample_indices.tolist() == [0] and got.hamming_dists.tolist() == [0]
    st.close()


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3", "C4"])
@pytest.mark.parametrize("kernel", KERNELS)
def test_baseline_configs_parity_sample(cfg, kernel):
    """BASELINE.json configs on a bounded sample the oracle finishes in seconds (C1 at its full 1M)."""
    w = sg.synth.CONFIGS[cfg]
    n = {"C1": 1_000_000, "C2": 2_000_000, "C3": 400_000, "C4": 2_000_000}[cfg]
    table = w.table()
    reads = w.reads_host(0, n, table)
    kind = sg.select_matcher(w.barcode_len)
    st = _stage(table, kind, w.max_mismatches, w.min_delta, w.free_ns, w.max_no_calls, w.index_lengths, kernel)
    got = st.demultiplex(reads)
    want = _oracle(table, kind, w.max_mismatches, w.min_delta, w.free_ns, w.max_no_calls, w.index_lengths, reads)
    _assert_same(st, got, want, n)
    if cfg == "C4":  # override: cached hamming on the short inline barcode + most_frequent_unmatched counts
        c = sg.demux.count_unmatched(got.unmatched)
        assert c.counts == orc.unmatched_counts(reads, want["idx"], w.num_samples)
        for mnc in (0, 1):
            s2 = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, mnc, None, kernel)
            g2 = s2.demultiplex(reads)
            _assert_same(s2, g2, _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, mnc, None, reads), n)
            s2.close()
    st.close()


def test_full_size_properties_c2():
    """At a large size the oracle cannot afford, check size-independent properties: conservation
    (sum of per-sample totals == n == total_templates), shard-merge invariance, idempotence, and
    agreement of the two kernels with each other."""
    import torch
    w = sg.synth.C2
    n = 20_000_000
    table = w.table()
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, w.barcode_len), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    outs = []
    for kernel in KERNELS:
        st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, kernel)
        d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
        d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
        st.matcher.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        st.matcher.sync(0)
        m1 = st.metrics()
        per = m1.per_sample_array()
        assert int(per[:, 0].sum()) == n == m1.total_templates
        hist = torch.bincount(d_idx.to(torch.int64) & 0xFFFF, minlength=w.num_samples + 1).cpu().numpy()
        assert np.array_equal(hist.astype(np.uint64), per[:, 0])
        assert int(((d_dist == 0) & (d_idx.to(torch.int64) < w.num_samples)).sum()) == int(per[:-1, 1].sum())
        assert 0.85 * n < int(per[:-1, 0].sum()) < 0.95 * n
        # shards on two slots accumulate to exactly twice the single pass
        half = n // 2
        st.matcher.match_device(d_in.data_ptr(), half, d_idx.data_ptr(), d_dist.data_ptr(), slot=0)
        st.matcher.match_device(d_in.data_ptr() + half * w.barcode_len, n - half, d_idx.data_ptr() + 2 * half,
                                d_dist.data_ptr() + half, slot=1)
        st.matcher.sync(0)
        st.matcher.sync(1)
        assert np.array_equal(st.metrics().per_sample_array(), per * 2)
        hops = st.metrics().sample_barcode_hop_tracker.count_tracker
        assert all(v % 2 == 0 for v in hops.values()) and 0.02 * n < sum(hops.values()) // 2 < 0.04 * n
        outs.append((d_idx.cpu().numpy().copy(), d_dist.cpu().numpy().copy(), per, hops))
        st.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]
    # the first 1M reads against the oracle
    k = 1_000_000
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, w.reads_host(0, k, table))
    assert np.array_equal(outs[0][0][:k], want["idx"]) and np.array_equal(outs[0][1][:k], want["dist"])


def test_batch_larger_than_4_gib_of_barcodes():
    """One 250 M-read C2 batch = 5 GB of ASCII barcodes: byte offsets need 64 bits.  Conservation, bincount of
    the index array == counters, and the first and the last million reads against the oracle."""
    import torch
    w = sg.synth.C2
    n = 250_000_000
    table = w.table()
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, w.barcode_len), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, sg.KERNEL_AUTO)
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    st.matcher.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
    st.matcher.sync(0)
    m = st.metrics()
    per = m.per_sample_array()
    assert m.total_templates == n == int(per[:, 0].sum())
    hist = torch.zeros(w.num_samples + 1, dtype=torch.int64, device="cuda")
    for a in range(0, n, 50_000_000):  # bincount in slices: the int64 copy of the whole array would be 2 GB
        hist += torch.bincount(d_idx[a:a + 50_000_000].to(torch.int64) & 0xFFFF, minlength=w.num_samples + 1)
    assert np.array_equal(hist.cpu().numpy().astype(np.uint64), per[:, 0])
    k = 1_000_000
    for first in (0, n - k):
        want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths,
                       w.reads_host(first, k, table))
        assert np.array_equal(d_idx[first:first + k].cpu().numpy().view(np.uint16), want["idx"])
        assert np.array_equal(d_dist[first:first + k].cpu().numpy(), want["dist"])
    st.close()


def test_staged_front_end_is_bit_exact(monkeypatch):
    """SGDX_STAGED=1 selects the bulk-copy (cp.async.bulk + mbarrier) front end of sgdx_match_exact_first, an
    experiment that is not the default: full tiles come through shared memory, the ragged
    last tile through ordinary loads.  Same verdicts and counters as the oracle."""
    import torch
    monkeypatch.setenv("SGDX_STAGED", "1")
    w = sg.synth.C2
    n = 1_000_003
    table = w.table()
    reads = w.reads_host(0, n, table)
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, w.max_mismatches, w.min_delta, w.free_ns, None,
                w.index_lengths, sg.KERNEL_AUTO)
    d_in = torch.from_numpy(reads).cuda()
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    st.matcher.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
    st.matcher.sync(0)
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, w.max_mismatches, w.min_delta, w.free_ns, None,
                   w.index_lengths, reads)
    assert np.array_equal(d_idx.cpu().numpy().view(np.uint16), want["idx"])
    assert np.array_equal(d_dist.cpu().numpy(), want["dist"])
    m = st.metrics()
    assert np.array_equal(m.per_sample_array(), want["per_sample"]) and m.total_templates == n
    assert m.sample_barcode_hop_tracker.count_tracker == want["hops"]
    st.close()


def test_full_size_properties_c3_with_unmatched_and_routing():
    """BASELINE config C3 at its full size (1536 x (12+12), M=3, delta=1, 500 M reads = 12 GB of barcodes) in one
    launch, with the unmatched table on and the routing after it: conservation, bincount of the index array ==
    counters == run lengths of the routing, every routed run sorted (stable) and holding only its sample, the
    unmatched totals == Undetermined's count, and the first / last half million reads against the oracle."""
    import torch
    w = sg.synth.C3
    n = w.n_reads
    assert n == 500_000_000
    table = w.table()
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, w.barcode_len), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    st = sg.BarcodeAssignmentStage(_samples(table), sg.CachedHammingDistanceMatcher, w.max_mismatches, w.min_delta,
                                   w.free_ns, None, w.index_lengths, unmatched_on_device=True,
                                   most_unmatched_max_map_size=60_000_000)
    S = w.num_samples
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_order = torch.empty(n, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
    st.matcher.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
    st.matcher.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr())
    st.matcher.sync(0)
    per = st.metrics().per_sample_array()
    assert int(per[:, 0].sum()) == n == st.metrics().total_templates
    hist = torch.zeros(S + 1, dtype=torch.int64, device="cuda")
    for a in range(0, n, 50_000_000):
        hist += torch.bincount(d_idx[a:a + 50_000_000].to(torch.int64) & 0xFFFF, minlength=S + 1)
    assert np.array_equal(hist.cpu().numpy().astype(np.uint64), per[:, 0])
    off = d_off.cpu().numpy()
    assert off[0] == 0 and off[-1] == n and np.array_equal(np.diff(off).astype(np.uint64), per[:, 0])
    for s in (0, 1, S // 2, S - 1, S):  # a few runs: ascending read ids (stable), all of that sample
        run = d_order[int(off[s]):int(off[s + 1])].to(torch.int64)
        if run.numel() > 1:
            assert bool((run[1:] > run[:-1]).all())
        assert bool(((d_idx[run].to(torch.int64) & 0xFFFF) == s).all())
    info = st.matcher.unmatched_info()
    assert info["unmatched_reads"] == int(per[S, 0]) and info["dropped_reads"] == 0 and info["downsizes"] == 0
    top = st.matcher.unmatched_top(100)
    assert len(top) == 100 and all(top[i][1] >= top[i + 1][1] for i in range(99))
    k = 500_000
    for first in (0, n - k):
        reads = w.reads_host(first, k, table)
        want = _oracle(table, sg.MatcherKind.CachedHammingDistance, w.max_mismatches, w.min_delta, w.free_ns, None,
                       w.index_lengths, reads)
        assert np.array_equal(d_idx[first:first + k].cpu().numpy().view(np.uint16), want["idx"])
        assert np.array_equal(d_dist[first:first + k].cpu().numpy(), want["dist"])
    um = orc.unmatched_counts(reads, want["idx"], S)  # the last slice's own unmatched keys are in the table
    full = dict(top)
    heavy = max(um.items(), key=lambda kv: kv[1])
    assert heavy[1] <= full.get(heavy[0], 10 ** 18)
    st.close()
