"""GPU parity of the host-packed 16-byte record path (sgdx_pack_host + sgdx_match_packed_batch: SURVEY Appendix B's
`sgdx_pack` / packed `sgdx_match_batch`) against the CPU oracle: index, distance, per-sample counters, hop map, for six
table shapes (8 to 32 bases) x both rule sets x three rule settings."""
import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
pytestmark = pytest.mark.gpu

ORC_KIND = {sg.MatcherKind.CachedHammingDistance: orc.MATCHER_CACHED_HAMMING, sg.MatcherKind.PreCompute: orc.MATCHER_PRECOMPUTE}
CLS = {sg.MatcherKind.CachedHammingDistance: sg.CachedHammingDistanceMatcher, sg.MatcherKind.PreCompute: sg.PreComputeMatcher}


@pytest.mark.parametrize("S,L,i1", [(1, 8, 0), (96, 16, 8), (384, 20, 10), (300, 24, 12), (50, 32, 16), (33, 13, 5)])
def test_host_packed_batches_match_oracle(S, L, i1):
    rng = np.random.default_rng(S * 31 + L)
    n = 20_000
    w = sg.synth.Workload("r", S, L, i1, 1, 2, n_reads=n, table_seed=int(rng.integers(1 << 30)),
                          read_seed=int(rng.integers(1 << 30)), p_true=0.80, p_hop=0.08 if i1 else 0.0,
                          p_sub=0.03, p_n=0.02, p_x=0.004, min_dist=2)
    table = w.table()
    reads = w.reads_host(0, n, table)
    il = w.index_lengths
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for (M, D, F, mnc) in [(1, 2, 1, None), (2, 1, 2, 1), (0, 0, 0, None)]:
            m = CLS[kind]([bytes(t) for t in table], M, D, F, max_no_calls=mnc, index_lengths=il)
            idx, dist = m.find_batch_packed(reads, threads=2)
            cfg = orc.Config(S, L, M, D, F, -1 if mnc is None else mnc, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
            want = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)
            assert np.array_equal(idx, want["idx"]) and np.array_equal(dist, want["dist"])
            per, totals = m.counters()
            assert np.array_equal(per, want["per_sample"]) and totals["total_templates"] == n
            if il:
                assert m.hop_counts() == want["hops"]
            m.close()


LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def _run_packed_device(m, reads, shift=0, with_dist=True):
    """16-byte records resident on the device (sgdx_match_packed_device), optionally moved off the vector alignment by
    `shift` records (records themselves stay 16-byte aligned, as the ABI requires)."""
    import torch
    n = len(reads)
    rec = sg.pack_host(reads) if n else np.zeros((0, 4), dtype=np.uint32)
    d_rec = torch.zeros((n + 4, 4), dtype=torch.int32, device="cuda")
    d_rec[shift:shift + n] = torch.from_numpy(rec.view(np.int32)).cuda()
    d_idx = torch.full((n + 8,), -1, dtype=torch.int16, device="cuda")
    d_dist = torch.full((n + 8,), 0x77, dtype=torch.uint8, device="cuda")
    m.match_packed_device(d_rec.data_ptr() + 16 * shift, n, d_idx.data_ptr() + 2 * shift, (d_dist.data_ptr() + shift) if with_dist else 0)
    m.sync()
    torch.cuda.synchronize()
    idx, dist = d_idx.cpu().numpy().view(np.uint16), d_dist.cpu().numpy()
    assert (idx[:shift] == 0xFFFF).all() and (idx[shift + n:] == 0xFFFF).all()
    assert (dist[:shift] == 0x77).all() and (dist[shift + n:] == 0x77).all()
    return idx[shift:shift + n], dist[shift:shift + n]


@pytest.mark.parametrize("n,shift", [(0, 0), (1, 0), (63, 1), (64, 0), (65, 1), (300_001, 0), (300_001, 1)])
def test_c3_table_runs_the_perfect_hash_kernel_on_16_byte_records(n, shift):
    """BASELINE config 3 (1536 x (12+12), -m 3 -d 1): 24 bases do not fit a compact record, so the kernel runs on
    sgdx_read records with the 1536 barcodes themselves as its key set; every other read goes through its general
    path (seed index in shared memory), N-heavy and foreign bytes included."""
    w0 = sg.synth.C3
    table = w0.table()
    S, L = table.shape
    w = sg.synth.Workload("t", S, L, w0.index1_len, 3, 1, table_seed=w0.table_seed, read_seed=5, p_true=0.85, p_hop=0.06,
                          p_sub=0.03, p_n=0.01, p_x=0.002)
    reads = w.reads_host(0, n, table) if n else np.zeros((0, L), dtype=np.uint8)
    if n > 1000:
        k = n // 50
        rng = np.random.default_rng(n)
        a, b = rng.integers(0, S, k), rng.integers(0, S, k)
        reads[:k, :12], reads[:k, 12:] = table[b, 12:], table[a, :12]   # reversed-orientation hops
        reads[k:k + 20, :12] = ord("N")
        reads[k + 20:k + 30, :] = ord("N")
    il = w.index_lengths
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for (M, D, F, mnc) in [(3, 1, 1, None), (1, 2, 1, None), (3, 1, 0, 2), (0, 0, 0, None)]:
            m = CLS[kind]([bytes(t) for t in table], M, D, F, max_no_calls=mnc, index_lengths=il)
            assert m.packed_kernel_name() == "sgdx_match_phx" and m.path_flags() & 64
            assert bool(m.path_flags() & 128) == (M >= 1)  # short seed index whenever the key set stops short of M
            idx, dist = _run_packed_device(m, reads, shift)
            per, totals = m.counters()
            assert totals["total_templates"] == n
            if n:
                cfg = orc.Config(S, L, M, D, F, -1 if mnc is None else mnc, ORC_KIND[kind], il[0], il[1])
                want = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)
                assert np.array_equal(idx, want["idx"]) and np.array_equal(dist, want["dist"])
                assert np.array_equal(per, want["per_sample"])
                assert m.hop_counts() == want["hops"]
            m.close()


@pytest.mark.parametrize("seed,S,L,i1,M,D", [(21, 50, 24, 12, 1, 1), (22, 20, 32, 16, 1, 2), (23, 8, 32, 0, 2, 0),
                                             (24, 100, 22, 10, 1, 3), (25, 300, 26, 13, 1, 1), (26, 40, 32, 16, 3, 1)])
def test_wide_records_on_close_tables(seed, S, L, i1, M, D):
    """22..32-base barcodes without a distance guarantee (ties, delta rejections, no single-N shortcut), all-T / all-G
    reads (a read of 32 T's has the planes of the kernel's dummy row), no distance output."""
    rng = np.random.default_rng(seed)
    table = np.unique(LETTERS[rng.integers(0, 4, (S * 2, L))], axis=0)
    rng.shuffle(table)
    table = np.ascontiguousarray(table[:S])
    for i in range(0, len(table) - 1, 5):
        table[i + 1] = table[i]
        table[i + 1, rng.integers(0, L)] = LETTERS[rng.integers(0, 4)]
    table = np.ascontiguousarray(np.unique(table, axis=0))
    S = len(table)
    n = 40_007
    reads = table[rng.integers(0, S, n)].copy()
    sub = rng.random(reads.shape) < 0.04
    reads[sub] = LETTERS[rng.integers(0, 4, int(sub.sum()))]
    reads[-4000:] = LETTERS[rng.integers(0, 4, (4000, L))]
    reads[rng.random(reads.shape) < 0.01] = ord("N")
    reads[rng.random(reads.shape) < 0.001] = ord("R")
    reads[0], reads[1], reads[2] = ord("T"), ord("G"), ord("A")
    il = (i1, L - i1) if i1 else None
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for (F, mnc) in ((1, None), (0, 1)):
            m = CLS[kind]([bytes(t) for t in table], M, D, F, max_no_calls=mnc, index_lengths=il)
            assert m.packed_kernel_name() == "sgdx_match_phx"
            idx, _ = _run_packed_device(m, reads, 0, with_dist=False)
            cfg = orc.Config(S, L, M, D, F, -1 if mnc is None else mnc, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
            want = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)
            assert np.array_equal(idx, want["idx"])
            per, totals = m.counters()
            assert np.array_equal(per, want["per_sample"]) and totals["total_templates"] == n
            if il:
                assert m.hop_counts() == want["hops"]
            m.close()


def test_c2_at_two_mismatches_uses_radius_one_and_the_seed_index():
    """384 x 20 bases at -m 2: the strings within two substitutions are too many, so compact records run the
    perfect-hash kernel with the radius-1 key set and its general path decides the rest."""
    w = sg.synth.C2
    table = w.table()
    n = 500_003
    import dataclasses
    reads = dataclasses.replace(w, p_sub=0.04, p_n=0.005).reads_host(0, n, table)
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for (M, D) in ((2, 1), (3, 2)):
            st = sg.BarcodeAssignmentStage([sg.SampleMetadata(f"s{i}", bytes(b), i) for i, b in enumerate(table)] +
                                           [sg.SampleMetadata(sg.UNDETERMINED_NAME, b"N" * 20, len(table))],
                                           CLS[kind], M, D, 1, None, w.index_lengths)
            assert st.matcher.compact_kernel_name() == "sgdx_match_phx"
            got = st.demultiplex(reads, compact=True, pack_threads=4)
            cfg = orc.Config(len(table), 20, M, D, 1, -1, ORC_KIND[kind], 10, 10)
            want = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)
            assert np.array_equal(got.sample_indices, want["idx"]) and np.array_equal(got.hamming_dists, want["dist"])
            met = st.metrics()
            assert np.array_equal(met.per_sample_array(), want["per_sample"])
            assert met.sample_barcode_hop_tracker.count_tracker == want["hops"]
            st.close()


@pytest.mark.parametrize("S,L,i1", [(300, 24, 12), (50, 32, 16), (64, 22, 11), (40, 29, 0), (1536, 24, 12)])
def test_dense_records_of_long_barcodes_give_the_same_results(S, L, i1):
    """22..32 bases as dense records (6..8 bytes per read over PCIe instead of 16) + an exception list of 16-byte records:
    the device rebuilds the sgdx_read records and runs the same kernels.  Oracle parity, ragged sizes, no exceptions,
    nothing but exceptions."""
    rng = np.random.default_rng(S * 17 + L)
    n = 50_003
    if S == 1536:
        w = sg.synth.C3
        table = w.table()
    else:
        w = sg.synth.Workload("r", S, L, i1, 1, 2, n_reads=n, table_seed=int(rng.integers(1 << 30)), read_seed=int(rng.integers(1 << 30)),
                              p_true=0.80, p_hop=0.08 if i1 else 0.0, p_sub=0.03, p_n=0.02, p_x=0.004, min_dist=2)
        table = w.table()
    import dataclasses
    reads = dataclasses.replace(w, p_n=0.02, p_x=0.004).reads_host(0, n, table)
    il = w.index_lengths
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        M, D = (3, 1) if S == 1536 else (1, 2)
        m = CLS[kind]([bytes(t) for t in table], M, D, 1, index_lengths=il)
        cfg = orc.Config(len(table), L, M, D, 1, -1, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
        want = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)
        idx, dist = m.find_batch_dense(reads, threads=2)
        assert np.array_equal(idx, want["idx"]) and np.array_equal(dist, want["dist"])
        per, totals = m.counters()
        assert np.array_equal(per, want["per_sample"]) and totals["total_templates"] == n
        if il:
            assert m.hop_counts() == want["hops"]
        for k in (0, 1, 3, 5, 4097):
            i2, d2 = m.find_batch_dense(reads[:k], slot=1)
            assert np.array_equal(i2, want["idx"][:k]) and np.array_equal(d2, want["dist"][:k])
        plain = np.isin(reads, LETTERS).all(axis=1)
        for sel in (plain, ~plain):
            part = np.ascontiguousarray(reads[sel][:3001])
            i3, d3 = m.find_batch_dense(part)
            assert np.array_equal(i3, want["idx"][sel][:3001]) and np.array_equal(d3, want["dist"][sel][:3001])
        m.close()
