"""GPU parity of the compact-record path (include/sgdx.h: sgdx_pack_compact_host / _device,
sgdx_match_compact_batch / _device; kernel sgdx_match_ph, or the general kernels on expanded records) against the CPU oracle and against the ASCII
path, bit-exact on index, distance, per-sample counters, total_templates and the hop map."""
import ctypes as C
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


def _random_case(rng, S, L, i1, n=20000):
    w = sg.synth.Workload("r", S, L, i1, 1, 2, n_reads=n, table_seed=int(rng.integers(1 << 30)),
                          read_seed=int(rng.integers(1 << 30)), p_true=0.80, p_hop=0.08 if i1 else 0.0,
                          p_sub=0.03, p_n=0.02, p_x=0.004, min_dist=1 if S > 4 ** min(L, 4) // 8 else 2)
    table = w.table()
    reads = w.reads_host(0, n, table)
    if i1:  # reversed-orientation hop rule and Undetermined's all-N halves
        k = n // 50
        a, b = rng.integers(0, S, k), rng.integers(0, S, k)
        if i1 * 2 == L:
            reads[:k, :i1], reads[:k, i1:] = table[b, i1:], table[a, :i1]
        reads[k:k + 20, :i1] = ord("N")
        reads[k + 20:k + 40, i1:] = ord("N")
        reads[k + 40:k + 50, :] = ord("N")
    return w, table, reads


PARAMS = [(1, 2, 1, None), (0, 0, 0, None), (2, 1, 2, 1), (3, 1, 1, 0), (1, 0, 0, 2), (2, 3, 1, None)]


@pytest.mark.parametrize("S,L,i1", [(1, 8, 0), (2, 4, 2), (7, 12, 0), (96, 16, 8), (97, 20, 10), (384, 20, 10),
                                    (200, 21, 10), (33, 13, 5), (64, 6, 3), (3, 1, 0)])
def test_compact_path_matches_oracle_on_random_tables(S, L, i1):
    """Perfect-hash kernel where the table's match neighbourhood is small enough, 16-byte general path elsewhere: both through
    sgdx_match_compact_batch, both rule sets, N-heavy and foreign bytes included."""
    rng = np.random.default_rng(S * 977 + L)
    w, table, reads = _random_case(rng, S, L, i1)
    il = w.index_lengths
    fast = 0
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for (M, D, F, mnc) in PARAMS:
            st = _stage(table, kind, M, D, F, mnc, il)
            fast += bool(st.matcher.path_flags() & PATH_PH)
            got = st.demultiplex(reads, compact=True, pack_threads=2)
            want = _oracle(table, kind, M, D, F, mnc, il, reads)
            _assert_same(st, got, want, reads.shape[0])
            st.close()
    if (S, L) in ((96, 16), (384, 20), (200, 21)):
        assert fast > 0, "the perfect-hash kernel never engaged on a table it should serve"


@pytest.mark.parametrize("L", [1, 7, 8, 16, 19, 20, 21])
def test_device_packer_writes_the_host_packers_records(L):
    import torch
    rng = np.random.default_rng(L)
    alphabet = np.frombuffer(b"ACGTACGTACGTNNacgn\x00\x01\x80\xff\x4f\x51", dtype=np.uint8)
    n = 50_001
    arr = alphabet[rng.integers(0, alphabet.size, size=(n, L))]
    m = sg.CachedHammingDistanceMatcher([b"ACGTACGTACGTACGTACGTA"[:L]], 1, 2, 1)
    d_in = torch.from_numpy(arr.copy()).cuda()
    d_rec = torch.empty(n, dtype=torch.int64, device="cuda")
    m.pack_compact_device(d_in.data_ptr(), n, d_rec.data_ptr())
    m.sync()
    torch.cuda.synchronize()
    assert np.array_equal(d_rec.cpu().numpy().view(np.uint64), sg.pack_compact_host(arr, threads=2))
    m.close()


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_baseline_tables_take_the_compact_kernel_and_equal_the_ascii_path(cfg):
    """BASELINE C1 / C2 tables, 3 M + 1 reads (many tiles per CTA, ragged tail): device-resident compact records
    against the ASCII kernel on the same reads -- index, distance, counters, hop map."""
    import torch
    w = getattr(sg.synth, cfg)
    n = 3_000_001
    table = w.table()
    reads = w.reads_host(0, n, table)
    out = {}
    for mode in ("ascii", "compact"):
        st = _stage(table, sg.MatcherKind.CachedHammingDistance, w.max_mismatches, w.min_delta, w.free_ns, None, w.index_lengths)
        m = st.matcher
        assert m.path_flags() & PATH_PH
        d_in = torch.from_numpy(reads).cuda()
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
        out[mode] = (d_idx.cpu().numpy().view(np.uint16), d_dist.cpu().numpy(), met.per_sample_array(), met.total_templates,
                     met.sample_barcode_hop_tracker.count_tracker)
        st.close()
    a, c = out["ascii"], out["compact"]
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) and np.array_equal(a[2], c[2])
    assert a[3] == c[3] == n and a[4] == c[4]
    # and a slice against the oracle
    k = 200_000
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, w.max_mismatches, w.min_delta, w.free_ns, None, w.index_lengths, reads[:k])
    assert np.array_equal(c[0][:k], want["idx"]) and np.array_equal(c[1][:k], want["dist"])


def test_compact_empty_tiny_and_refused_inputs():  # (unmatched table from records: tests/test_gpu_post.py)
    table = sg.synth.C2.table()
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, (10, 10))
    for n in (0, 1, 31, 2049):
        reads = np.ascontiguousarray(np.tile(table[:7], (n // 7 + 1, 1))[:n])
        got = st.demultiplex(reads, compact=True)
        assert got.sample_indices.tolist() == [i % 7 for i in range(n)]
        assert got.hamming_dists.tolist() == [0] * n
    st.close()
    long = sg.CachedHammingDistanceMatcher([b"ACGT" * 6, b"TTGA" * 6], 1, 2, 1)
    rec = np.zeros(4, dtype=np.uint64)
    idx = np.zeros(4, dtype=np.uint16)
    with pytest.raises(sg.SgdxError, match="at most 21"):
        long.submit_compact_ptr(rec.ctypes.data, 4, idx.ctypes.data, 0)
    long.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_batcher_in_compact_mode_returns_the_same_batches(mode):
    """sgdx_batcher_set_compact: push packs into the pinned buffer, batches go up as 8 bytes per read (mode 1) or as dense
    records + exception list (mode 2)."""
    from singular_demux_b200._lib import check, lib
    rng = np.random.default_rng(3)
    w, table, reads = _random_case(rng, 96, 16, 8, n=100_000)
    want = _oracle(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, reads)
    st = _stage(table, sg.MatcherKind.CachedHammingDistance, 1, 2, 1, None, w.index_lengths, slots=2)
    b = C.c_void_p()
    check(lib().sgdx_batcher_create(st.matcher.handle, 16_384, C.byref(b)))
    check(lib().sgdx_batcher_set_compact(b, mode))
    got_idx, got_dist = [], []

    def take():
        pi, pd, n = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
        check(lib().sgdx_batcher_next(b, C.byref(pi), C.byref(pd), C.byref(n)))
        if n.value:
            got_idx.append(np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint16)), (n.value,)).copy())
            got_dist.append(np.ctypeslib.as_array(C.cast(pd, C.POINTER(C.c_uint8)), (n.value,)).copy())
        return n.value

    refused = 0
    for chunk in np.array_split(reads, 100):  # 1000-read chunks like --chunksize
        c = np.ascontiguousarray(chunk)
        off = 0
        while off < c.shape[0]:  # 2 x slots batches may wait for sgdx_batcher_next: then the push stops and says how far it got
            took = C.c_uint64(0)
            rc = lib().sgdx_batcher_push_n(b, c.ctypes.data + off * c.shape[1], c.shape[0] - off, C.byref(took))
            off += took.value
            if rc == -4:
                refused += 1
                assert take() > 0
            else:
                check(rc)
    assert refused >= 1
    assert lib().sgdx_batcher_set_compact(b, 0) != 0  # refused while batches are in flight
    while lib().sgdx_batcher_flush(b) == -4:
        assert take() > 0
    while take():
        pass
    lib().sgdx_batcher_destroy(b)
    assert np.array_equal(np.concatenate(got_idx), want["idx"])
    assert np.array_equal(np.concatenate(got_dist), want["dist"])
    assert np.array_equal(st.metrics().per_sample_array(), want["per_sample"])
    st.close()


@pytest.mark.parametrize("S,L,i1", [(96, 16, 8), (384, 20, 10), (1, 8, 0), (40, 21, 10), (33, 13, 5), (4, 4, 0)])
def test_dense_records_give_the_same_results(S, L, i1):
    """sgdx_pack_dense_host + sgdx_match_dense_batch: 2 bits per base over PCIe, the reads with a no-call or a foreign
    byte as an exception list; the device rebuilds the compact records and runs the same kernels.  Oracle parity on
    index, distance, counters and hop map; ragged sizes (the unpack kernel takes four reads per thread), no exceptions,
    nothing but exceptions."""
    rng = np.random.default_rng(S * 53 + L)
    w, table, reads = _random_case(rng, S, L, i1, n=60_003)
    il = w.index_lengths
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        want = _oracle(table, kind, 1, 2, 1, None, il, reads)
        st = _stage(table, kind, 1, 2, 1, None, il)
        idx, dist = st.matcher.find_batch_dense(reads, threads=2)
        assert np.array_equal(idx, want["idx"]) and np.array_equal(dist, want["dist"])
        m = st.metrics()
        assert np.array_equal(m.per_sample_array(), want["per_sample"]) and m.total_templates == len(reads)
        if il:
            assert m.sample_barcode_hop_tracker.count_tracker == want["hops"]
        # ragged and tiny batches, on the other slot
        for n in (0, 1, 2, 3, 5, 4097):
            part = reads[:n]
            i2, d2 = st.matcher.find_batch_dense(part, slot=1)
            assert np.array_equal(i2, want["idx"][:n]) and np.array_equal(d2, want["dist"][:n])
        plain = np.isin(reads, np.frombuffer(b"ACGT", dtype=np.uint8)).all(axis=1)
        for sel in (plain, ~plain):  # no exception at all / every read an exception
            part = np.ascontiguousarray(reads[sel][:5001])
            i3, d3 = st.matcher.find_batch_dense(part)
            assert np.array_equal(i3, want["idx"][sel][:5001]) and np.array_equal(d3, want["dist"][sel][:5001])
        st.close()
    long = sg.CachedHammingDistanceMatcher([b"ACGT" * 10, b"TTGA" * 10], 1, 2, 1)  # 40 bases: ASCII entry points only
    with pytest.raises(sg.SgdxError):
        long.submit_dense_ptr(0, 0, 0, 0, 0, 0, 0)
    long.close()
