"""GPU parity of the perfect-hash kernel (sgdx_match_ph, csrc/sgdx_ph.cu; SGDX_PATH_PERFECT_HASH) -- through the C ABI
(sgdx_match_compact_batch / _device), bit-exact against the CPU oracle on index, distance, per-sample counters,
total_templates and the hop map; and against the older compact kernel (SGDX_DISABLE_PH=1) on the same reads."""
import os

import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
pytestmark = pytest.mark.gpu

ORC_KIND = {sg.MatcherKind.CachedHammingDistance: orc.MATCHER_CACHED_HAMMING, sg.MatcherKind.PreCompute: orc.MATCHER_PRECOMPUTE}
CLS = {sg.MatcherKind.CachedHammingDistance: sg.CachedHammingDistanceMatcher, sg.MatcherKind.PreCompute: sg.PreComputeMatcher}
PATH_PH = 64
LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)
HAM, PRE = sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute


def _samples(table):
    out = [sg.SampleMetadata(f"s{i}", bytes(b), i) for i, b in enumerate(table)]
    out.append(sg.SampleMetadata(sg.UNDETERMINED_NAME, b"N" * len(out[0].barcode), len(out)))
    return out


def _stage(table, kind, M, D, F, mnc, il, kernel=0, slots=2):
    return sg.BarcodeAssignmentStage(_samples(table), CLS[kind], M, D, F, mnc, il, kernel=kernel, slots=slots)


def _oracle(table, kind, M, D, F, mnc, il, reads):
    S, L = len(table), len(table[0])
    cfg = orc.Config(S, L, M, D, F, -1 if mnc is None else mnc, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
    return orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)


def _assert_same(stage, got, want, n):
    assert np.array_equal(got.sample_indices, want["idx"])
    assert np.array_equal(got.hamming_dists, want["dist"])
    m = stage.metrics()
    assert np.array_equal(m.per_sample_array(), want["per_sample"])
    assert m.total_templates == want["total_templates"] == n
    if stage.index_lengths:
        assert m.sample_barcode_hop_tracker.count_tracker == want["hops"]


def _reads(w, table, n, rng, n_rate=None):
    reads = w.reads_host(0, n, table)
    S, L = table.shape
    i1 = w.index1_len
    if i1:  # reversed-orientation hop rule and Undetermined's all-N halves
        k = n // 50
        a, b = rng.integers(0, S, k), rng.integers(0, S, k)
        if i1 * 2 == L:
            reads[:k, :i1], reads[:k, i1:] = table[b, i1:], table[a, :i1]
        reads[k:k + 20, :i1] = ord("N")
        reads[k + 20:k + 40, i1:] = ord("N")
        reads[k + 40:k + 50, :] = ord("N")
    if n_rate:
        reads[rng.random(reads.shape) < n_rate] = ord("N")
    return reads


RULES = [(1, 2, 1, None), (1, 2, 0, None), (1, 2, 2, 0), (0, 0, 0, None), (1, 0, 1, 1), (1, 5, 1, None)]


@pytest.mark.parametrize("cfg_name", ["C1", "C2", "C4"])
def test_baseline_tables_run_the_perfect_hash_kernel_and_equal_the_oracle(cfg_name):
    w0 = sg.synth.CONFIGS[cfg_name]
    table = w0.table()
    rng = np.random.default_rng(len(table))
    w = sg.synth.Workload("t", w0.num_samples, w0.barcode_len, w0.index1_len, 1, 2, table_seed=w0.table_seed,
                          read_seed=77, p_true=0.8, p_hop=0.08 if w0.index1_len else 0.0, p_sub=0.02, p_n=0.01, p_x=0.002)
    reads = _reads(w, table, 60_013, rng)
    for kind in (HAM, PRE):
        for (M, D, F, mnc) in RULES:
            st = _stage(table, kind, M, D, F, mnc, w.index_lengths)
            assert st.matcher.path_flags() & PATH_PH, (kind, M, D)
            got = st.demultiplex(reads, compact=True, pack_threads=2)
            _assert_same(st, got, _oracle(table, kind, M, D, F, mnc, w.index_lengths, reads), len(reads))
            st.close()


@pytest.mark.parametrize("seed,S,L,i1,M,D", [(1, 40, 6, 3, 1, 1), (2, 40, 6, 0, 2, 1), (3, 64, 8, 4, 1, 2), (4, 64, 8, 0, 2, 0),
                                             (5, 12, 5, 0, 1, 0), (6, 30, 7, 0, 3, 1), (7, 200, 10, 5, 1, 3), (8, 16, 4, 2, 4, 1),
                                             (9, 100, 12, 6, 1, 1), (10, 3, 21, 10, 1, 2), (11, 500, 16, 8, 1, 1)])
def test_close_tables_ties_and_delta(seed, S, L, i1, M, D):
    """Random barcodes without a distance guarantee: precomputed NoMatch verdicts, no single-N shortcut, reads with
    invalid bases through the seed search or the all-samples scan; N-heavy stream."""
    rng = np.random.default_rng(seed)
    table = np.unique(LETTERS[rng.integers(0, 4, (S * 2, L))], axis=0)
    rng.shuffle(table)
    table = np.ascontiguousarray(table[:S])
    S = len(table)
    n = 30_001
    reads = table[rng.integers(0, S, n)].copy()
    sub = rng.random(reads.shape) < 0.06
    reads[sub] = LETTERS[rng.integers(0, 4, int(sub.sum()))]
    reads[-3000:] = LETTERS[rng.integers(0, 4, (3000, L))]
    reads[rng.random(reads.shape) < 0.02] = ord("N")
    reads[rng.random(reads.shape) < 0.002] = ord("R")
    il = (i1, L - i1) if i1 else None
    for kind in (HAM, PRE):
        for (F, mnc) in ((1, None), (0, 1), (2, None)):
            st = _stage(table, kind, M, D, F, mnc, il)
            if not st.matcher.path_flags() & PATH_PH:
                st.close()
                pytest.skip("too many keys for a perfect hash")
            got = st.demultiplex(reads, compact=True, pack_threads=2)
            _assert_same(st, got, _oracle(table, kind, M, D, F, mnc, il, reads), n)
            st.close()


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("n_rate", [0.02, 0.15, 0.5])
def test_no_call_heavy_reads(kernel, n_rate):
    """The general path of the kernel: single-N shortcut (C2's table allows it), several invalid bases through the
    seed index (kernel AUTO) or, with the direct kernel selected (no seed index), through the all-samples scan."""
    w = sg.synth.C2
    table = w.table()
    rng = np.random.default_rng(int(n_rate * 100) + kernel)
    reads = _reads(w, table, 40_000, rng, n_rate=n_rate)
    for kind in (HAM, PRE):
        for (M, D, F, mnc) in ((1, 2, 1, None), (1, 2, 3, 2), (1, 3, 0, None)):
            st = _stage(table, kind, M, D, F, mnc, w.index_lengths, kernel=kernel)
            assert st.matcher.path_flags() & PATH_PH
            got = st.demultiplex(reads, compact=True, pack_threads=2)
            _assert_same(st, got, _oracle(table, kind, M, D, F, mnc, w.index_lengths, reads), len(reads))
            st.close()


@pytest.mark.parametrize("n", [0, 1, 31, 127, 128, 129, 4095, 4097, 1_000_003])
@pytest.mark.parametrize("shift", [0, 1])
def test_ragged_sizes_and_unaligned_device_buffers(n, shift):
    """Device-resident records; shift = 1 moves every buffer off its vector alignment (64-bit loads, 16/8-bit stores)."""
    import torch
    w = sg.synth.C2
    table = w.table()
    reads = w.reads_host(5, n, table) if n else np.zeros((0, 20), dtype=np.uint8)
    rec = sg.pack_compact_host(reads) if n else np.zeros(0, dtype=np.uint64)
    st = _stage(table, HAM, 1, 2, 1, None, w.index_lengths)
    d_rec = torch.zeros(n + 8, dtype=torch.int64, device="cuda")
    d_idx = torch.full((n + 8,), -1, dtype=torch.int16, device="cuda")
    d_dist = torch.full((n + 8,), 0x77, dtype=torch.uint8, device="cuda")
    d_rec[shift:shift + n] = torch.from_numpy(rec.view(np.int64)).cuda()
    st.matcher.match_compact_device(d_rec.data_ptr() + 8 * shift, n, d_idx.data_ptr() + 2 * shift, d_dist.data_ptr() + shift)
    st.matcher.sync()
    torch.cuda.synchronize()
    idx = d_idx.cpu().numpy().view(np.uint16)
    dist = d_dist.cpu().numpy()
    # nothing outside [shift, shift + n) was touched
    assert (idx[:shift] == 0xFFFF).all() and (idx[shift + n:] == 0xFFFF).all()
    assert (dist[:shift] == 0x77).all() and (dist[shift + n:] == 0x77).all()
    if n:
        want = _oracle(table, HAM, 1, 2, 1, None, w.index_lengths, reads)
        assert np.array_equal(idx[shift:shift + n], want["idx"])
        assert np.array_equal(dist[shift:shift + n], want["dist"])
        m = st.metrics()
        assert np.array_equal(m.per_sample_array(), want["per_sample"])
        assert m.sample_barcode_hop_tracker.count_tracker == want["hops"]
    st.close()


def test_no_distance_output_and_stray_bits_are_ignored():
    """out_dist = NULL; bits above barcode_len in any plane and bit 63 do not change a record's meaning."""
    import torch
    w = sg.synth.C1  # 16 bases: 5 spare bits per plane
    table = w.table()
    n = 50_000
    rng = np.random.default_rng(5)
    reads = _reads(w, table, n, rng, n_rate=0.01)
    rec = sg.pack_compact_host(reads)
    L = 16
    stray = np.zeros(n, dtype=np.uint64)
    for plane in range(3):
        bits = rng.integers(0, 32, n).astype(np.uint64) * (rng.random(n) < 0.3)
        stray |= (bits << np.uint64(L)) << np.uint64(21 * plane)
    stray |= (rng.random(n) < 0.2).astype(np.uint64) << np.uint64(63)
    want = _oracle(table, HAM, 1, 2, 1, None, w.index_lengths, reads)
    for disable in ("", "1"):
        os.environ["SGDX_DISABLE_PH"] = disable
        if not disable:
            del os.environ["SGDX_DISABLE_PH"]
        try:
            st = _stage(table, HAM, 1, 2, 1, None, w.index_lengths)
        finally:
            os.environ.pop("SGDX_DISABLE_PH", None)
        assert bool(st.matcher.path_flags() & PATH_PH) == (disable == "")
        d_rec = torch.from_numpy((rec | stray).view(np.int64)).cuda()
        d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
        st.matcher.match_compact_device(d_rec.data_ptr(), n, d_idx.data_ptr(), 0)
        st.matcher.sync()
        torch.cuda.synchronize()
        assert np.array_equal(d_idx.cpu().numpy().view(np.uint16), want["idx"])
        assert np.array_equal(st.metrics().per_sample_array(), want["per_sample"])
        assert st.metrics().sample_barcode_hop_tracker.count_tracker == want["hops"]
        st.close()


def test_full_size_properties_and_agreement_with_the_older_kernels():
    """C2 at 20 M reads: conservation, bincount of the index array == counters, and the same index / distance arrays,
    counters and hop map as the older compact kernel and the ASCII kernel on the same reads."""
    import torch
    w = sg.synth.C2
    table = w.table()
    n = 20_000_000
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, 20), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), 0)
    torch.cuda.synchronize()
    res = {}
    for mode in ("ph", "compact", "ascii"):
        if mode == "compact":
            os.environ["SGDX_DISABLE_PH"] = "1"
        try:
            st = _stage(table, HAM, 1, 2, 1, None, w.index_lengths)
        finally:
            os.environ.pop("SGDX_DISABLE_PH", None)
        m = st.matcher
        assert bool(m.path_flags() & PATH_PH) == (mode != "compact")
        d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
        d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
        if mode == "ascii":
            m.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        else:
            d_rec = torch.empty(n, dtype=torch.int64, device="cuda")
            m.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr())
            m.match_compact_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        m.sync()
        torch.cuda.synchronize()
        met = st.metrics()
        per = met.per_sample_array()
        assert met.total_templates == n == int(per[:, 0].sum())
        counts = torch.bincount(d_idx.to(torch.int64) & 0xFFFF, minlength=len(table) + 1).cpu().numpy()
        assert np.array_equal(counts, per[:, 0].astype(np.int64))
        res[mode] = (d_idx.cpu(), d_dist.cpu(), per, met.sample_barcode_hop_tracker.count_tracker)
        st.close()
        del d_idx, d_dist
    for other in ("compact", "ascii"):
        assert torch.equal(res["ph"][0], res[other][0]) and torch.equal(res["ph"][1], res[other][1])
        assert np.array_equal(res["ph"][2], res[other][2]) and res["ph"][3] == res[other][3]
    k = 300_000
    reads = d_in[:k].cpu().numpy()
    want = _oracle(table, HAM, 1, 2, 1, None, w.index_lengths, reads)
    assert np.array_equal(res["ph"][0][:k].numpy().view(np.uint16), want["idx"])
    assert np.array_equal(res["ph"][1][:k].numpy(), want["dist"])


@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_ragged_batch_above_the_one_launch_threshold(cfg):
    """A batch of more than 32 M reads whose size is not a multiple of the chunk: the vector kernel takes the full
    chunks and a second launch the last reads (smaller ragged batches take the general kernel alone).  Same arrays,
    counters and hop map as the ASCII kernel over all reads; the oracle on both ends."""
    import torch
    w = sg.synth.CONFIGS[cfg]
    table = w.table()
    L = w.barcode_len
    n = (32 << 20) + 4096 + 77
    d_table = torch.from_numpy(table).cuda()
    d_in = torch.empty((n, L), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), 0)
    d_in[n - 40:n - 30] = ord("N")
    torch.cuda.synchronize()
    res = {}
    for mode in ("records", "ascii"):
        st = _stage(table, HAM, w.max_mismatches, w.min_delta, 1, None, w.index_lengths)
        m = st.matcher
        d_idx = torch.full((n,), -1, dtype=torch.int16, device="cuda")
        d_dist = torch.full((n,), 0x77, dtype=torch.uint8, device="cuda")
        if mode == "ascii":
            m.match_device(d_in.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        elif L <= 21:
            d_rec = torch.empty(n, dtype=torch.int64, device="cuda")
            m.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr())
            m.match_compact_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        else:
            d_rec = torch.empty((n, 4), dtype=torch.int32, device="cuda")
            m.pack_device(d_in.data_ptr(), n, d_rec.data_ptr())
            m.match_packed_device(d_rec.data_ptr(), n, d_idx.data_ptr(), d_dist.data_ptr())
        m.sync()
        torch.cuda.synchronize()
        met = st.metrics()
        assert met.total_templates == n
        res[mode] = (d_idx.cpu(), d_dist.cpu(), met.per_sample_array(), met.sample_barcode_hop_tracker.count_tracker)
        st.close()
        del d_idx, d_dist
    assert torch.equal(res["records"][0], res["ascii"][0]) and torch.equal(res["records"][1], res["ascii"][1])
    assert np.array_equal(res["records"][2], res["ascii"][2]) and res["records"][3] == res["ascii"][3]
    for sl in (slice(0, 100_000), slice(n - 100_000, n)):
        want = _oracle(table, HAM, w.max_mismatches, w.min_delta, 1, None, w.index_lengths, d_in[sl].cpu().numpy())
        assert np.array_equal(res["records"][0][sl].numpy().view(np.uint16), want["idx"])
        assert np.array_equal(res["records"][1][sl].numpy(), want["dist"])


def test_stress_streams_equal_the_oracle():
    """bench.py's "stress" streams on C2's table: 5 % substitutions per base (a third of the reads end up one
    mismatch away, a third unmatched) and exactly one substituted base in every read."""
    import dataclasses
    w = sg.synth.C2
    table = w.table()
    n = 400_003
    rng = np.random.default_rng(11)
    a = dataclasses.replace(w, p_sub=0.05).reads_host(0, n, table)
    b = dataclasses.replace(w, p_true=1.0, p_hop=0.0, p_sub=0.0, p_n=0.0).reads_host(0, n, table)
    pos = rng.integers(0, 20, n)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    cur = np.vectorize(code.get)(b[np.arange(n), pos])
    b[np.arange(n), pos] = LETTERS[(cur + rng.integers(1, 4, n)) & 3]
    for reads in (a, b):
        st = _stage(table, HAM, 1, 2, 1, None, w.index_lengths)
        got = st.demultiplex(reads, compact=True, pack_threads=2)
        _assert_same(st, got, _oracle(table, HAM, 1, 2, 1, None, w.index_lengths, reads), n)
        st.close()
    assert (got.hamming_dists == 1).all() and (got.sample_indices < len(table)).all()


@pytest.mark.parametrize("n,shift", [(0, 0), (1, 0), (3, 1), (4, 0), (5, 0), (1027, 1), (400_001, 0), (400_001, 1)])
def test_single_sample_kernel_sizes_alignment_and_agreement_with_the_perfect_hash(n, shift):
    """BASELINE config 4 (one 8-base inline barcode, 2 % no-calls per base, PreCompute by run.rs:196): compact records
    run sgdx_match_single; ragged sizes, unaligned buffers, both rule sets with and without the no-call gate, against
    the oracle and against sgdx_match_ph on the same records (SGDX_DISABLE_SINGLE=1)."""
    import torch
    w = sg.synth.C4
    table = w.table()
    rng = np.random.default_rng(n + shift)
    reads = dataclass_reads(w, table, n, rng)
    rec = sg.pack_compact_host(reads) if n else np.zeros(0, dtype=np.uint64)
    for kind in (PRE, HAM):
        for (M, D, F, mnc) in ((1, 2, 1, None), (2, 0, 0, 1), (0, 3, 2, 0)):
            want = _oracle(table, kind, M, D, F, mnc, None, reads) if n else None
            outs = []
            for disable in ("", "1"):
                os.environ["SGDX_DISABLE_SINGLE"] = disable
                if not disable:
                    del os.environ["SGDX_DISABLE_SINGLE"]
                try:
                    st = _stage(table, kind, M, D, F, mnc, None)
                    assert st.matcher.compact_kernel_name() == ("sgdx_match_single" if not disable else "sgdx_match_ph")
                    d_rec = torch.zeros(n + 8, dtype=torch.int64, device="cuda")
                    d_idx = torch.full((n + 8,), -1, dtype=torch.int16, device="cuda")
                    d_dist = torch.full((n + 8,), 0x77, dtype=torch.uint8, device="cuda")
                    d_rec[shift:shift + n] = torch.from_numpy(rec.view(np.int64)).cuda()
                    st.matcher.match_compact_device(d_rec.data_ptr() + 8 * shift, n, d_idx.data_ptr() + 2 * shift, d_dist.data_ptr() + shift)
                    st.matcher.sync()
                    torch.cuda.synchronize()
                finally:
                    os.environ.pop("SGDX_DISABLE_SINGLE", None)
                idx, dist = d_idx.cpu().numpy().view(np.uint16), d_dist.cpu().numpy()
                assert (idx[:shift] == 0xFFFF).all() and (idx[shift + n:] == 0xFFFF).all()
                assert (dist[:shift] == 0x77).all() and (dist[shift + n:] == 0x77).all()
                m = st.metrics()
                outs.append((idx[shift:shift + n].copy(), dist[shift:shift + n].copy(), m.per_sample_array(), m.total_templates))
                st.close()
            a, b = outs
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3] == n
            if n:
                assert np.array_equal(a[0], want["idx"]) and np.array_equal(a[1], want["dist"])
                assert np.array_equal(a[2], want["per_sample"])


def dataclass_reads(w, table, n, rng):
    """C4-shaped reads with some foreign bytes and a few all-N / heavily masked reads."""
    if n == 0:
        return np.zeros((0, w.barcode_len), dtype=np.uint8)
    import dataclasses
    reads = dataclasses.replace(w, p_x=0.003).reads_host(3, n, table)
    k = min(n, 50)
    reads[:k][rng.random((k, w.barcode_len)) < 0.5] = ord("N")
    return reads
