"""GPU parity tests of the stages either side of the assignment: most-frequent unmatched
barcodes, segment extraction, base-quality counters, read routing -- all through the C ABI, bit-exact
against oracle/oracle.py on the same seeded inputs.  Run on a B200: python -m pytest tests -m gpu."""
import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
pytestmark = pytest.mark.gpu

RS = sg.ReadStructure.from_str
ORC_KIND = {sg.MatcherKind.CachedHammingDistance: orc.MATCHER_CACHED_HAMMING, sg.MatcherKind.PreCompute: orc.MATCHER_PRECOMPUTE}
CLS = {sg.MatcherKind.CachedHammingDistance: sg.CachedHammingDistanceMatcher, sg.MatcherKind.PreCompute: sg.PreComputeMatcher}


def _samples(table):
    out = [sg.SampleMetadata(f"s{i}", bytes(b), i) for i, b in enumerate(table)]
    out.append(sg.SampleMetadata(sg.UNDETERMINED_NAME, b"N" * len(out[0].barcode), len(out)))
    return out


def _stage(w, table, kind=None, **kw):
    kind = kind or sg.select_matcher(w.barcode_len)
    return sg.BarcodeAssignmentStage(_samples(table), CLS[kind], w.max_mismatches, w.min_delta, w.free_ns,
                                     w.max_no_calls, w.index_lengths, **kw)


def _oracle_idx(w, table, reads, kind=None):
    kind = kind or sg.select_matcher(w.barcode_len)
    il = w.index_lengths
    cfg = orc.Config(w.num_samples, w.barcode_len, w.max_mismatches, w.min_delta, w.free_ns,
                     -1 if w.max_no_calls is None else w.max_no_calls, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
    return orc.COracle(cfg, [bytes(t) for t in table]).run(reads, threads=8, use_cache=True)["idx"]


def _segs(rs):
    return [(g.offset, g.length, g.kind) for g in rs.segments]


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------
# 1. unmatched barcodes
# ------------------------------------------------------------------------------------------------
def test_unmatched_known_answers_on_device(golden):
    """run.rs:652-689 (three unmatched keys, once each), run.rs:1382-1451 (TTT+AAA), run.rs:1455-1605."""
    seen = 0
    for case in golden["demux_cases"]:
        if "unmatched" not in case:
            continue
        table = [s.encode() for s in case["samples"]]
        L = len(table[0])
        reads = np.frombuffer("".join(case["reads"]).encode(), dtype=np.uint8).reshape(-1, L)
        il = (case["index1_len"], case["index2_len"]) if case["index1_len"] else None
        for k in case["matchers"]:
            kind = sg.MatcherKind.PreCompute if k == 1 else sg.MatcherKind.CachedHammingDistance
            st = sg.BarcodeAssignmentStage(_samples(table), CLS[kind], case["max_mismatches"], case["min_delta"],
                                           case["free_ns"], case["max_no_calls"], il, unmatched_on_device=True)
            st.demultiplex(reads)
            got = {k_.decode(): v for k_, v in st.matcher.unmatched_counts().items()}
            assert got == case["unmatched"], case["src"]
            rows = st.most_frequent_unmatched(1000)
            assert sorted(r.count for r in rows) == sorted(case["unmatched"].values())
            if il and case["unmatched"]:
                assert all(r.barcode[il[0]] == "+" for r in rows)
            info = st.matcher.unmatched_info()
            assert info["unmatched_reads"] == sum(case["unmatched"].values()) and info["dropped_reads"] == 0
            st.close()
            seen += 1
    assert seen >= 4


@pytest.mark.parametrize("cfg,n", [("C1", 300_000), ("C2", 400_000), ("C3", 200_000), ("C4", 400_000)])
def test_unmatched_counts_match_oracle(cfg, n):
    """Exact counts of every unmatched barcode on the BASELINE configs (C4: 2 % no-calls, PreCompute rule),
    fed in three batches on two slots; the top-N and its threshold selection against a host sort."""
    w = sg.synth.CONFIGS[cfg]
    table = w.table()
    reads = w.reads_host(0, n, table)
    if cfg == "C4":  # a few foreign bytes: those keys are counted on the host side of the library
        reads[::997, 3] = ord("R")
        reads[5::4001, 0] = ord("n")
    st = _stage(w, table, unmatched_on_device=True, slots=2)
    cuts = [0, n // 3, n // 2, n]
    idx = np.concatenate([st.demultiplex(reads[a:b], slot=i % 2).sample_indices
                          for i, (a, b) in enumerate(zip(cuts, cuts[1:]))])
    want_idx = _oracle_idx(w, table, reads)
    assert np.array_equal(idx, want_idx)
    want = orc.unmatched_counts(reads, want_idx, w.num_samples)
    got = st.matcher.unmatched_counts()
    assert got == want
    info = st.matcher.unmatched_info()
    assert info == {"distinct_keys": len(want), "unmatched_reads": int((want_idx == w.num_samples).sum()),
                    "dropped_reads": 0, "downsizes": 0}
    lit = orc.UnmatchedCounterLiteral()
    for r in reads[want_idx == w.num_samples]:
        lit.insert(r.tobytes())
    assert lit.unmatched_counter == got
    for topn in (1, 10, 1000, len(want), len(want) + 5):
        rows = st.matcher.unmatched_top(topn)
        ref = lit.top(topn)
        assert len(rows) == len(ref) == min(topn, len(want))
        assert [c for _, c in rows] == [c for _, c in ref]           # same counts, descending
        assert all(want[k] == c for k, c in rows)                     # every reported key has its true count
        cut = ref[-1][1] if ref else 0
        assert {k for k, c in ref if c > cut} <= {k for k, _ in rows}  # everything above the tie is present
    st.matcher.reset_counters()
    assert st.matcher.unmatched_counts() == {}
    st.close()


def test_unmatched_downsize_follows_the_reference_heuristic():
    """metrics.rs:151-175 at batch granularity: once max_map_size distinct keys are reached the table is cut to
    the downsize_to most frequent.  Heavy keys survive with exact counts; totals are conserved up to the cut."""
    w = sg.synth.Workload("d", 24, 16, 8, 1, 2, n_reads=0, table_seed=77, read_seed=78, p_true=0.5, p_hop=0.0)
    table = w.table()
    rng = np.random.default_rng(5)
    heavy = rng.integers(0, 4, (50, 16))
    heavy = np.frombuffer(b"ACGT", dtype=np.uint8)[heavy]
    st = _stage(w, table, unmatched_on_device=True, most_unmatched_max_map_size=3000, most_unmatched_downsize_to=100)
    lit_heavy = {}
    for b in range(6):
        noise = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (1500, 16))]
        reps = np.repeat(heavy, 40 + b, axis=0)
        batch = np.concatenate([noise, reps])
        rng.shuffle(batch)
        idx = st.demultiplex(batch).sample_indices
        um = batch[idx == w.num_samples]
        for r in um:
            k = r.tobytes()
            lit_heavy[k] = lit_heavy.get(k, 0) + 1
    info = st.matcher.unmatched_info()
    assert info["downsizes"] >= 1 and info["dropped_reads"] == 0
    assert info["distinct_keys"] < 3000 + 1500 + 60
    got = st.matcher.unmatched_counts()
    heavy_keys = [k for k, c in lit_heavy.items() if c >= 40]
    assert len(heavy_keys) >= 40
    for k in heavy_keys:           # far above the noise: never cut, counted exactly
        assert got[k] == lit_heavy[k]
    top = st.matcher.unmatched_top(len(heavy_keys))
    assert {k for k, _ in top} == set(heavy_keys)
    st.matcher.unmatched_downsize()
    assert st.matcher.unmatched_info()["distinct_keys"] <= 100
    assert {k for k, _ in st.matcher.unmatched_top(len(heavy_keys))} == set(heavy_keys)
    st.close()


def test_unmatched_long_barcodes_and_all_n():
    """L = 32 uses both key words (21 + 11 bases); all-N reads are the classic top unmatched key."""
    w = sg.synth.Workload("l", 48, 32, 16, 2, 1, n_reads=0, table_seed=9, read_seed=10, p_true=0.6, p_hop=0.1, p_n=0.01)
    table = w.table()
    reads = w.reads_host(0, 120_000, table)
    reads[::50] = ord("N")
    st = _stage(w, table, unmatched_on_device=True)
    idx = st.demultiplex(reads).sample_indices
    want_idx = _oracle_idx(w, table, reads)
    assert np.array_equal(idx, want_idx)
    want = orc.unmatched_counts(reads, want_idx, w.num_samples)
    assert st.matcher.unmatched_counts() == want
    assert st.matcher.unmatched_top(1) == [(b"N" * 32, 2400)]
    assert st.most_frequent_unmatched(1)[0] == sg.BarcodeCount("N" * 16 + "+" + "N" * 16, 2400)
    st.close()


def test_unmatched_enable_state_checks():
    w = sg.synth.C1
    table = w.table()
    st = _stage(w, table)
    with pytest.raises(sg.SgdxError):
        st.matcher.unmatched_counts()          # not enabled
    st.matcher.unmatched_enable(1000, 10)
    with pytest.raises(sg.SgdxError):
        st.matcher.unmatched_enable(1000, 10)  # twice
    st.close()


@pytest.mark.parametrize("cfg,n,form", [("C1", 300_000, "compact"), ("C2", 400_000, "compact"), ("C4", 400_000, "compact"),
                                        ("C3", 200_000, "packed"), ("C2", 200_000, "packed"), ("C2m2", 200_000, "compact"),
                                        ("C2m4", 60_000, "compact")])
def test_unmatched_counts_from_compact_and_packed_records(cfg, n, form):
    """Config 4's most_frequent_unmatched on the fast input formats: the unmatched table is keyed from the records
    (compact 64-bit records through sgdx_match_ph / sgdx_match_phx / sgdx_match_single, 16-byte records through
    sgdx_match_phx or the general kernels) and holds exactly the oracle's counts; a NoMatch barcode with a byte that is
    none of A/C/G/T/N has no key in a record and is reported as dropped."""
    import dataclasses
    base = sg.synth.CONFIGS[cfg[:2]]
    w = base if len(cfg) == 2 else dataclasses.replace(base, max_mismatches=int(cfg[3]), min_delta=1)
    table = w.table()
    reads = dataclasses.replace(w, p_n=0.01).reads_host(0, n, table)
    reads[::50] = ord("N")
    reads[7::997, 3] = ord("R")
    st = _stage(w, table, unmatched_on_device=True, slots=2)
    cuts = [0, n // 3, n // 2, n]
    parts = []
    for i, (a, b) in enumerate(zip(cuts, cuts[1:])):
        if form == "compact":
            parts.append(st.demultiplex(reads[a:b], slot=i % 2, compact=True, pack_threads=2).sample_indices)
        else:
            parts.append(st.matcher.find_batch_packed(reads[a:b], slot=i % 2, threads=2)[0])
    idx = np.concatenate(parts)
    want_idx = _oracle_idx(w, table, reads)
    assert np.array_equal(idx, want_idx)
    um = want_idx == w.num_samples
    foreign = (~np.isin(reads, np.frombuffer(b"ACGTN", dtype=np.uint8))).any(axis=1)
    keyable = um & ~foreign
    want = orc.unmatched_counts(reads[keyable], want_idx[keyable], w.num_samples)
    got = st.matcher.unmatched_counts()
    assert got == want
    info = st.matcher.unmatched_info()
    assert info["unmatched_reads"] == int(um.sum()) and info["dropped_reads"] == int((um & foreign).sum())
    assert info["distinct_keys"] == len(want)
    assert (b"N" * w.barcode_len) in got
    st.close()


# ------------------------------------------------------------------------------------------------
# 2. extraction
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("structs,strides,L", [
    (["10B", "10B"], [10, 10], 20),                      # two index reads
    (["8B142T", "150T", "8B"], [150, 150, 8], 16),       # inline + index
    (["4S8B+T"], [151], 8),                              # skip, inline barcode, template
    (["3B2S4B+T", "2M5B"], [40, 7], 12),                 # several B segments per read, odd length
    (["+B"], [9], 9),
])
def test_extraction_matches_oracle_and_feeds_the_matcher(structs, strides, L):
    import torch
    rng = np.random.default_rng(L * 1000 + len(structs))
    n = 70_001
    S = 40
    w = sg.synth.Workload("x", S, L, 0, 1, 1, n_reads=n, table_seed=3, read_seed=4, p_n=0.01)
    table = w.table()
    bc = w.reads_host(0, n, table)
    rss = [RS(s) for s in structs]
    srcs, pos = [], 0
    for rs, stride in zip(rss, strides):  # scatter the synthetic barcodes into their segments
        a = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, (n, stride))].copy()
        for g in rs.segments:
            if g.kind == "B":
                ln = (stride - g.offset) if g.length is None else g.length
                a[:, g.offset:g.offset + ln] = bc[:, pos:pos + ln]
                pos += ln
        srcs.append(a)
    assert pos == L
    want = orc.extract_sample_barcodes(srcs, [_segs(rs) for rs in rss])
    assert np.array_equal(want, bc)
    m = sg.CachedHammingDistanceMatcher(list(table), 1, 1, 1)
    d_srcs = [_cuda(a) for a in srcs]
    for shift in (0, 1):  # 4-byte aligned and unaligned output
        d_out = torch.zeros(n * L + 8, dtype=torch.uint8, device="cuda")
        m.extract_device([d.data_ptr() for d in d_srcs], strides, rss, n, d_out.data_ptr() + shift)
        m.sync(0)
        assert np.array_equal(d_out[shift:shift + n * L].cpu().numpy().reshape(n, L), want)
        assert int(d_out[shift + n * L:].sum()) == 0 and int(d_out[:shift].sum()) == 0
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    m.match_segments_device([d.data_ptr() for d in d_srcs], strides, rss, n, d_idx.data_ptr(), d_dist.data_ptr())
    m.sync(0)
    idx2, dist2 = m.find_batch(bc, slot=1)
    assert np.array_equal(d_idx.cpu().numpy().view(np.uint16), idx2)
    assert np.array_equal(d_dist.cpu().numpy(), dist2)
    m.close()


def test_extraction_rejects_wrong_structures():
    import torch
    m = sg.CachedHammingDistanceMatcher([b"ACGTACGT", b"TTTTGGGG"], 1, 1, 1)
    d = torch.zeros(1000, dtype=torch.uint8, device="cuda")
    o = torch.zeros(1000, dtype=torch.uint8, device="cuda")
    for structs, strides in [(["7B"], [7]), (["9B"], [9]), (["8T"], [8]), (["8B"], [6])]:
        with pytest.raises(sg.SgdxError):
            m.extract_device([d.data_ptr()], strides, [RS(s) for s in structs], 10, o.data_ptr())
    m.close()


# ------------------------------------------------------------------------------------------------
# 3. base-quality counters
# ------------------------------------------------------------------------------------------------
def _quals(rng, n, stride, bad=False):
    q = rng.integers(33, 75, (n, stride), dtype=np.uint8)
    q[rng.random((n, stride)) < 0.3] = ord("I")
    if bad:  # bytes a FASTQ should not hold: below '!' (u8 wrap in the reference) and above '~'
        q[rng.random((n, stride)) < 0.002] = 10
        q[rng.random((n, stride)) < 0.002] = 200
        q[rng.random((n, stride)) < 0.002] = 32
        q[rng.random((n, stride)) < 0.002] = 128
    return q


def test_base_qual_known_answer_on_device():
    """run.rs:1040-1157: 17B39T, qualities '?'*17 + 35..73 -> per read 39 bases, 11 >= Q30, 21 >= Q20."""
    q = np.array([[63] * 17 + list(range(35, 74))] * 4, dtype=np.uint8)
    m = sg.CachedHammingDistanceMatcher([b"A" * 17, b"C" * 17, b"G" * 17, b"T" * 17], 2, 1, 1)
    d_q, d_idx = _cuda(q), _cuda(np.zeros(4, dtype=np.uint16))
    m.qual_counts_device(d_q.data_ptr(), 4, 56, RS("17B39T"), d_idx.data_ptr())
    got, short = m.base_qual_counters()
    assert got[0].tolist() == [84, 44, 17 * 30 * 4, 17 * 4, 156] and short == 0
    assert int(got[1:].sum()) == 0
    c = sg.BaseQualCounter.from_row(got[0])
    assert (c.q20_bases, c.q30_bases, c.bases) == (84, 44, 156)
    m.close()


@pytest.mark.parametrize("struct,stride,S,ragged,bad", [
    ("150T", 150, 384, False, False),        # C2's template reads
    ("+T", 151, 384, True, False),           # ragged lengths, odd stride
    ("8B+T", 158, 1, True, False),           # C4 inline
    ("10B", 10, 96, False, False),           # index read
    ("17B39T", 56, 4, False, True),          # reference's test structure, invalid bytes
    ("6M+T", 40, 16, True, True),
    ("3S5B2M20T4S+T", 64, 7, True, False),
    ("300T", 300, 1536, False, False),
    ("+T", 16, 3, True, True),
    ("2500T", 2500, 5, False, False),        # tiles of 16 reads
    ("12B", 12, 8191, False, False),         # histogram too large for shared memory: global atomics
])
def test_base_qual_counts_match_oracle(struct, stride, S, ragged, bad):
    rng = np.random.default_rng(stride * 7 + S)
    n = 50_003 if stride <= 300 else 3_001
    rs = RS(struct)
    q = _quals(rng, n, stride, bad)
    idx = rng.integers(0, S + 1, n).astype(np.uint16)
    idx[rng.random(n) < 0.5] = 0
    lens = None
    if ragged:
        lens = rng.integers(max(rs.min_length() - 2, 0), stride + 1, n).astype(np.uint32)
    want, want_short = orc.base_qual_counts(q, lens, _segs(rs), idx, S)
    L = 4
    m = sg.CachedHammingDistanceMatcher([bytes(b) for b in sg.synth.Workload("q", S, 8, 0, 1, 1, table_seed=1, min_dist=1).table()], 1, 1, 1)
    d_q, d_idx = _cuda(q), _cuda(idx)
    d_lens = _cuda(lens) if lens is not None else None
    half = (n // 2) & ~15  # two calls on two slots: the counters are cumulative
    m.qual_counts_device(d_q.data_ptr(), half, stride, rs, d_idx.data_ptr(), d_lens.data_ptr() if ragged else 0, slot=0)
    m.qual_counts_device(d_q.data_ptr() + half * stride, n - half, stride, rs, d_idx.data_ptr() + 2 * half,
                         d_lens.data_ptr() + 4 * half if ragged else 0, slot=1)
    got, short = m.base_qual_counters()
    assert short == want_short
    assert np.array_equal(got, want)
    if ragged:
        assert want_short > 0 or rs.min_length() <= 2
    m.reset_counters()
    assert int(m.base_qual_counters()[0].sum()) == 0
    m.close()


def test_base_qual_two_fastqs_of_one_template():
    """demux.rs:504-520: the counters of every FASTQ of a template go to the template's sample."""
    rng = np.random.default_rng(11)
    n, S = 30_000, 12
    idx = rng.integers(0, S + 1, n).astype(np.uint16)
    q1, q2, qi = _quals(rng, n, 100), _quals(rng, n, 75), _quals(rng, n, 8)
    m = sg.CachedHammingDistanceMatcher([bytes(b) for b in sg.synth.Workload("q", S, 8, 0, 1, 1, table_seed=1).table()], 1, 1, 1)
    d_idx = _cuda(idx)
    want = np.zeros((S + 1, 5), dtype=np.uint64)
    keep = []
    for q, s in ((q1, "100T"), (q2, "5M+T"), (qi, "8B")):
        d = _cuda(q)
        keep.append(d)
        m.qual_counts_device(d.data_ptr(), n, q.shape[1], RS(s), d_idx.data_ptr())
        want += orc.base_qual_counts(q, None, _segs(RS(s)), idx, S)[0]
    got, _ = m.base_qual_counters()
    assert np.array_equal(got, want)
    assert int(got[:, 4].sum()) == n * 170 and int(got[:, 3].sum()) == n * 8
    m.close()


def test_base_qual_rejects_bad_arguments():
    import torch
    m = sg.CachedHammingDistanceMatcher([b"ACGTACGT"], 1, 1, 1)
    d = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    i = torch.zeros(16, dtype=torch.int16, device="cuda")
    with pytest.raises(sg.SgdxError):
        m.qual_counts_device(d.data_ptr() + 1, 4, 100, RS("100T"), i.data_ptr())     # unaligned
    with pytest.raises(sg.SgdxError):
        m.qual_counts_device(d.data_ptr(), 4, 50, RS("100T"), i.data_ptr())          # structure longer than stride
    with pytest.raises(sg.SgdxError):
        m.qual_counts_device(d.data_ptr(), 4, 4000, RS("+T"), i.data_ptr())          # stride too large
    m.close()


# ------------------------------------------------------------------------------------------------
# 4. routing
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,n", [(1, 0), (1, 1), (3, 31), (3, 33), (96, 100_003), (384, 1_000_000), (1536, 250_000),
                                 (8192, 40_000), (384, 5_000_000), (2200, 300_000), (3000, 300_000)])
def test_routing_is_a_stable_partition(S, n):
    import torch
    rng = np.random.default_rng(S + n)
    idx = (rng.random(n) ** 2 * (S + 1)).astype(np.uint16)
    idx[rng.random(n) < 0.08] = S
    m = sg.CachedHammingDistanceMatcher([bytes(b) for b in sg.synth.Workload("r", S, 12, 0, 1, 1, table_seed=1, min_dist=1).table()], 1, 1, 1)
    d_idx = _cuda(idx) if n else torch.zeros(1, dtype=torch.int16, device="cuda")
    d_order = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
    for _ in range(2):  # scratch reuse
        m.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr())
        m.sync(0)
    want_order, want_off = orc.route_reads(idx, S)
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want_off)
    if n:
        assert np.array_equal(d_order.cpu().numpy().view(np.uint32), want_order)
    m.close()


@pytest.mark.parametrize("S,kind", [(2, "skew"), (3, "runs"), (4, "skew"), (384, "skew"), (384, "runs"), (384, "one"),
                                    (1536, "skew"), (40, "unaligned")])
def test_routing_of_skewed_index_arrays(S, kind):
    """Steps of the ranked scatter dominated by one bin (the rounds give up and the hardware match takes over), runs of
    equal indices, a single bin, and an index array that is not 16-byte aligned (scalar histogram loads)."""
    import torch
    n = 300_007
    rng = np.random.default_rng(S * 7 + len(kind))
    idx = rng.integers(0, S + 1, n).astype(np.uint16)
    if kind == "skew":
        idx[rng.random(n) < 0.9] = S // 2
    elif kind == "runs":
        idx = np.sort(idx)[::-1].copy()
        idx[1000:5000] = rng.integers(0, S + 1, 4000)
    elif kind == "one":
        idx[:] = S
    m = sg.CachedHammingDistanceMatcher([bytes(b) for b in sg.synth.Workload("r", S, 12, 0, 1, 1, table_seed=1, min_dist=1).table()], 1, 1, 1)
    off_bytes = 2 if kind == "unaligned" else 0
    buf = torch.zeros(n + 8, dtype=torch.int16, device="cuda")
    d_idx = buf[off_bytes // 2: off_bytes // 2 + n]
    d_idx.copy_(torch.from_numpy(idx.view(np.int16)))
    d_order = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
    m.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr())
    m.sync(0)
    want_order, want_off = orc.route_reads(idx, S)
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want_off)
    assert np.array_equal(d_order.cpu().numpy().view(np.uint32), want_order)
    m.close()


@pytest.mark.parametrize("S,n,kind", [(1, 4099, "flat"), (3, 70_001, "skew"), (96, 513, "flat"), (384, 1_000_003, "flat"),
                                      (384, 300_007, "skew"), (384, 200_000, "runs"), (700, 300_000, "flat"),
                                      (2200, 300_000, "flat"), (3000, 100_000, "flat")])
def test_routing_staged_form_on_small_batches(S, n, kind, monkeypatch):
    """The scatter that large batches take (votes for the ranking, turns between the warps of a CTA, whole sectors
    through a ring per bin), forced onto small batches through its hook: stable partition against the oracle for flat,
    skewed and sorted index arrays, with `order` at every word offset inside a 32-byte sector (the sector phase `a`),
    nothing written outside order[0..n); 3000 bins do not fit its rings and take the per-warp form."""
    import torch
    monkeypatch.setenv("SGDX_ROUTE_STAGED_FROM", "0")
    rng = np.random.default_rng(S * 11 + n)
    idx = rng.integers(0, S + 1, n).astype(np.uint16)
    if kind == "skew":
        idx[rng.random(n) < 0.9] = S
    elif kind == "runs":
        idx = np.sort(idx)[::-1].copy()
        idx[1000:5000] = rng.integers(0, S + 1, 4000)
    want_order, want_off = orc.route_reads(idx, S)
    m = sg.CachedHammingDistanceMatcher([bytes(b) for b in sg.synth.Workload("r", S, 12, 0, 1, 1, table_seed=1, min_dist=1).table()], 1, 1, 1)
    d_idx = _cuda(idx)
    for k in (0, 1, 3, 7):
        buf = torch.full((n + 16,), -1, dtype=torch.int32, device="cuda")
        d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
        m.route_device(d_idx.data_ptr(), n, buf.data_ptr() + 4 * k, d_off.data_ptr())
        m.sync(0)
        got = buf.cpu().numpy()
        assert (got[:k] == -1).all() and (got[k + n:] == -1).all()
        assert np.array_equal(got[k:k + n].view(np.uint32), want_order)
        assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want_off)
    m.close()


def test_routing_of_a_large_batch_takes_the_staged_form():
    """70 M reads (above the 64 M switch) without any hook: stable partition against numpy's stable argsort."""
    import torch
    S, n = 384, 70_000_037
    rng = np.random.default_rng(70)
    idx = rng.integers(0, S + 1, n, dtype=np.uint16)
    idx[::11] = S
    m = sg.CachedHammingDistanceMatcher([bytes(b) for b in sg.synth.Workload("r", S, 12, 0, 1, 1, table_seed=1, min_dist=1).table()], 1, 1, 1)
    d_idx = _cuda(idx)
    d_order = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
    m.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr())
    m.sync(0)
    want = np.argsort(idx, kind="stable").astype(np.uint32)
    assert np.array_equal(d_order.cpu().numpy().view(np.uint32), want)
    counts = np.bincount(idx, minlength=S + 1)
    assert np.array_equal(d_off.cpu().numpy()[:S + 2].astype(np.int64), np.concatenate([[0], np.cumsum(counts)]))
    m.close()


def test_routing_after_assignment_c2():
    """Assignment -> routing on one stream, device resident: every sample's run holds exactly its reads, in order."""
    import torch
    w = sg.synth.C2
    n = 2_000_000
    table = w.table()
    d_table = _cuda(table)
    d_in = torch.empty((n, w.barcode_len), dtype=torch.uint8, device="cuda")
    w.reads_device(0, n, d_table.data_ptr(), d_in.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    st = _stage(w, table)
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_order = torch.empty(n, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(w.num_samples + 2, dtype=torch.int64, device="cuda")
    st.matcher.match_device(d_in.data_ptr(), n, d_idx.data_ptr())
    st.matcher.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr())
    st.matcher.sync(0)
    idx = d_idx.cpu().numpy().view(np.uint16)
    want_order, want_off = orc.route_reads(idx, w.num_samples)
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want_off)
    assert np.array_equal(d_order.cpu().numpy().view(np.uint32), want_order)
    per = st.metrics().per_sample_array()
    assert np.array_equal(np.diff(want_off.astype(np.int64)), per[:, 0].astype(np.int64))
    st.close()


# ------------------------------------------------------------------------------------------------
# host threads, slots and streams
# ------------------------------------------------------------------------------------------------
def test_two_host_threads_share_one_handle_with_every_stage_on():
    """`M: Matcher + Send + Sync` (demux.rs:422): two host threads, one slot each, assignment + unmatched table +
    quality counters + routing on their own shards; the totals must equal the single-pass oracle."""
    import threading
    import torch
    w = sg.synth.C1
    n = 600_000
    table = w.table()
    reads = w.reads_host(0, n, table)
    rng = np.random.default_rng(3)
    quals = rng.integers(33, 75, (n, 48), dtype=np.uint8)
    rs = RS("8B+T")
    st = _stage(w, table, unmatched_on_device=True, slots=2)
    want_idx = _oracle_idx(w, table, reads)
    d_reads, d_quals = _cuda(reads), _cuda(quals)
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_order = torch.empty(n, dtype=torch.int32, device="cuda")
    d_off = torch.zeros((2, w.num_samples + 2), dtype=torch.int64, device="cuda")
    half = (n // 2) & ~15
    errors = []

    def work(slot, a, b):
        try:
            step = 37_504  # multiple of 16: quality tiles stay 16-byte aligned
            for lo in range(a, b, step):
                hi = min(lo + step, b)
                m = st.matcher
                m.match_device(d_reads.data_ptr() + lo * w.barcode_len, hi - lo, d_idx.data_ptr() + 2 * lo, 0, slot)
                m.qual_counts_device(d_quals.data_ptr() + lo * 48, hi - lo, 48, rs, d_idx.data_ptr() + 2 * lo, 0, slot)
                m.sync(slot)
            st.matcher.route_device(d_idx.data_ptr() + 2 * a, b - a, d_order.data_ptr() + 4 * a,
                                    d_off[slot].data_ptr(), slot)
            st.matcher.sync(slot)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(0, 0, half)), threading.Thread(target=work, args=(1, half, n))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    idx = d_idx.cpu().numpy().view(np.uint16)
    assert np.array_equal(idx, want_idx)
    assert st.matcher.unmatched_counts() == orc.unmatched_counts(reads, want_idx, w.num_samples)
    got_q, _ = st.matcher.base_qual_counters()
    want_q, _ = orc.base_qual_counts(quals, None, _segs(rs), want_idx, w.num_samples)
    assert np.array_equal(got_q, want_q)
    for slot, (a, b) in enumerate(((0, half), (half, n))):
        o, off = orc.route_reads(want_idx[a:b], w.num_samples)
        assert np.array_equal(d_order[a:b].cpu().numpy().view(np.uint32), o)
        assert np.array_equal(d_off[slot].cpu().numpy().astype(np.uint64), off)
    st.close()


def test_empty_batches_and_tiny_inputs_on_every_stage():
    """n = 0 and n = 1 through every entry point (ragged ends of a run; demux.rs handles empty chunks too)."""
    import torch
    w = sg.synth.C1
    table = w.table()
    st = _stage(w, table, unmatched_on_device=True)
    m = st.matcher
    d = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    i16 = torch.zeros(64, dtype=torch.int16, device="cuda")
    i32 = torch.zeros(64, dtype=torch.int32, device="cuda")
    off = torch.full((w.num_samples + 2,), -1, dtype=torch.int64, device="cuda")
    rs = RS("16B")
    m.match_device(d.data_ptr(), 0, i16.data_ptr())
    m.extract_device([d.data_ptr()], [16], [rs], 0, d.data_ptr() + 2048)
    m.match_segments_device([d.data_ptr()], [16], [rs], 0, i16.data_ptr())
    m.qual_counts_device(d.data_ptr(), 0, 16, rs, i16.data_ptr())
    m.route_device(i16.data_ptr(), 0, i32.data_ptr(), off.data_ptr())
    m.sync(0)
    assert off.cpu().numpy().tolist() == [0] * (w.num_samples + 2)
    assert m.unmatched_counts() == {} and m.unmatched_top(5) == []
    assert int(m.base_qual_counters()[0].sum()) == 0
    one = np.frombuffer(b"N" * 16, dtype=np.uint8).reshape(1, 16)
    q1 = np.full((1, 16), ord("I"), dtype=np.uint8)
    d_one, d_q1 = _cuda(one), _cuda(q1)
    m.match_segments_device([d_one.data_ptr()], [16], [rs], 1, i16.data_ptr())
    m.qual_counts_device(d_q1.data_ptr(), 1, 16, rs, i16.data_ptr())
    m.route_device(i16.data_ptr(), 1, i32.data_ptr(), off.data_ptr())
    m.sync(0)
    assert int(i16[0]) == w.num_samples and int(i32[0]) == 0
    assert off.cpu().numpy().tolist() == [0] * (w.num_samples + 1) + [1]
    assert m.unmatched_top(5) == [(b"N" * 16, 1)]
    got, short = m.base_qual_counters()
    assert got[w.num_samples].tolist() == [0, 0, 16 * 40, 16, 0] and short == 0
    info = m.unmatched_info()
    assert info["unmatched_reads"] == 1 and info["distinct_keys"] == 1
    st.close()


# ------------------------------------------------------------------------------------------------
# 5. one call per batch
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["dual_index_paired_end", "inline_single_end"])
def test_demultiplex_device_equals_the_per_batch_oracle(layout):
    """sgdx_demultiplex_device = Demultiplex::demultiplex for a device-resident batch (demux.rs:445-584): index,
    distance, per-sample counters, hop map, unmatched map, quality counters and per-sample read lists, all against
    oracle.demultiplex_batch (itself checked against the read-by-read restatement on CPU)."""
    import torch
    rng = np.random.default_rng(21)
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    if layout == "dual_index_paired_end":
        w = sg.synth.C2
        n = 150_000
        table = w.table()
        bc = w.reads_host(0, n, table)
        i1 = bc[:, :10].copy()
        i2 = letters[rng.integers(0, 4, (n, 12))].copy()
        i2[:, 2:12] = bc[:, 10:]
        r1 = letters[rng.integers(0, 5, (n, 151))].copy()
        r2 = letters[rng.integers(0, 5, (n, 100))].copy()
        arrays, structs = [i1, i2, r1, r2], ["10B", "2S10B", "+T", "8M92T"]
        ragged = {2: rng.integers(0, 152, n).astype(np.uint32)}
    else:
        w = sg.synth.C4
        n = 200_000
        table = w.table()
        bc = w.reads_host(0, n, table)
        r1 = letters[rng.integers(0, 5, (n, 158))].copy()
        r1[:, :8] = bc
        arrays, structs = [r1], ["8B+T"]
        ragged = {0: rng.integers(8, 159, n).astype(np.uint32)}
    rss = [RS(s) for s in structs]
    quals = [rng.integers(33, 75, a.shape, dtype=np.uint8) for a in arrays]
    kind = sg.select_matcher(w.barcode_len)
    il = w.index_lengths
    cfg = orc.Config(w.num_samples, w.barcode_len, w.max_mismatches, w.min_delta, w.free_ns, -1, ORC_KIND[kind],
                     il[0] if il else 0, il[1] if il else 0)
    want = orc.demultiplex_batch(cfg, [bytes(t) for t in table],
                                 [(a, q, ragged.get(i)) for i, (a, q) in enumerate(zip(arrays, quals))],
                                 [_segs(rs) for rs in rss])
    st = _stage(w, table, unmatched_on_device=True)
    m = st.matcher
    d_a, d_q = [_cuda(a) for a in arrays], [_cuda(q) for q in quals]
    d_l = {i: _cuda(v) for i, v in ragged.items()}
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    d_dist = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_order = torch.empty(n, dtype=torch.int32, device="cuda")
    d_off = torch.zeros(w.num_samples + 2, dtype=torch.int64, device="cuda")
    fastqs = [(d_a[i].data_ptr(), d_q[i].data_ptr(), d_l[i].data_ptr() if i in d_l else 0, arrays[i].shape[1])
              for i in range(len(arrays))]
    m.demultiplex_device(fastqs, rss, n, d_idx.data_ptr(), d_dist.data_ptr(), d_order.data_ptr(), d_off.data_ptr())
    m.sync(0)
    assert np.array_equal(d_idx.cpu().numpy().view(np.uint16), want["idx"])
    assert np.array_equal(d_dist.cpu().numpy(), want["dist"])
    met = st.metrics()
    assert np.array_equal(met.per_sample_array(), want["per_sample"]) and met.total_templates == n
    if il:
        assert met.sample_barcode_hop_tracker.count_tracker == want["hops"]
    assert m.unmatched_counts() == want["unmatched"]
    got_q, short = m.base_qual_counters()
    assert np.array_equal(got_q, want["base_qual"]) and short == 0
    assert np.array_equal(d_order.cpu().numpy().view(np.uint32), want["order"])
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want["offsets"])
    st.close()


def test_unmatched_table_under_heavy_contention():
    """Ten million reads, almost all of them unmatched: 40 % one key (all N), 30 % sixteen hot keys, the rest
    distinct random 24-mers -- same-slot claims and same-address adds from every CTA at once.  Exact counts."""
    rng = np.random.default_rng(99)
    w = sg.synth.Workload("h", 32, 24, 12, 1, 2, n_reads=0, table_seed=41, read_seed=42)
    table = w.table()
    n = 10_000_000
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = acgt[rng.integers(0, 4, (n, 24), dtype=np.uint8)]
    kind = rng.random(n)
    reads[kind < 0.4] = ord("N")
    hot = acgt[rng.integers(0, 4, (16, 24))]
    sel = np.flatnonzero((kind >= 0.4) & (kind < 0.7))
    reads[sel] = hot[rng.integers(0, 16, sel.size)]
    st = _stage(w, table, unmatched_on_device=True, slots=2, most_unmatched_max_map_size=8_000_000)
    half = n // 2
    import torch
    d_reads = _cuda(reads)
    d_idx = torch.empty(n, dtype=torch.int16, device="cuda")
    st.matcher.match_device(d_reads.data_ptr(), half, d_idx.data_ptr(), 0, 0)
    st.matcher.match_device(d_reads.data_ptr() + half * 24, n - half, d_idx.data_ptr() + 2 * half, 0, 1)
    st.matcher.sync(0)
    st.matcher.sync(1)
    idx = d_idx.cpu().numpy().view(np.uint16)
    um = reads[idx == w.num_samples]
    packed = np.ascontiguousarray(um).view(np.dtype((np.void, 24))).ravel()
    keys, counts = np.unique(packed, return_counts=True)
    info = st.matcher.unmatched_info()
    assert info["unmatched_reads"] == um.shape[0] and info["dropped_reads"] == 0 and info["downsizes"] == 0
    assert info["distinct_keys"] == keys.shape[0]
    top = st.matcher.unmatched_top(17)
    want_top = sorted(((k.tobytes(), int(c)) for k, c in zip(keys, counts) if c > 1), key=lambda kv: (-kv[1], kv[0]))[:17]
    assert top == want_top and top[0] == (b"N" * 24, int((kind < 0.4).sum()))
    got = st.matcher.unmatched_counts()
    assert len(got) == keys.shape[0]
    assert sum(got.values()) == um.shape[0]
    for k, c in zip(keys[::997], counts[::997]):
        assert got[k.tobytes()] == int(c)
    st.close()


def test_reference_extraction_tests_on_device():
    """demux.rs:867-1043 through sgdx_demultiplex_device, both matchers: every layout's template lands in sample 1,
    the extracted barcode is SAMPLE_BARCODE_1, `bases` = the template length, 17 index bases of quality 0."""
    import json
    import os
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = json.load(open(os.path.join(root, "tests", "golden", "reference_extraction_cases.json")))
    table = [s.encode() for s in g["samples"]]
    for case in g["cases"]:
        rss = [RS(s) for s in case["structures"]]
        for cls in (sg.CachedHammingDistanceMatcher, sg.PreComputeMatcher):
            st = sg.BarcodeAssignmentStage(_samples(table), cls, g["max_mismatches"], g["min_delta"], g["free_ns"],
                                           g["max_no_calls"], None, unmatched_on_device=True)
            m = st.matcher
            keep, fastqs = [], []
            for b in case["bases"]:
                a = torch.zeros(len(b) + 32, dtype=torch.uint8, device="cuda")  # 16-byte aligned, padded
                a[:len(b)] = torch.frombuffer(bytearray(b.encode()), dtype=torch.uint8).cuda()
                q = torch.full((len(b) + 32,), g["qual_byte"], dtype=torch.uint8, device="cuda")
                keep += [a, q]
                fastqs.append((a.data_ptr(), q.data_ptr(), 0, len(b)))
            d_bc = torch.zeros(32, dtype=torch.uint8, device="cuda")
            m.extract_device([f[0] for f in fastqs], [f[3] for f in fastqs], rss, 1, d_bc.data_ptr())
            d_idx = torch.full((4,), -1, dtype=torch.int16, device="cuda")
            d_dist = torch.full((4,), 77, dtype=torch.uint8, device="cuda")
            d_order = torch.full((4,), -1, dtype=torch.int32, device="cuda")
            d_off = torch.zeros(6, dtype=torch.int64, device="cuda")
            m.demultiplex_device(fastqs, rss, 1, d_idx.data_ptr(), d_dist.data_ptr(), d_order.data_ptr(), d_off.data_ptr())
            m.sync(0)
            assert bytes(d_bc[:17].cpu().numpy()) == table[0], case["src"]
            assert int(d_idx[0]) == g["expected_sample_index"] and int(d_dist[0]) == 0, case["src"]
            assert d_off.cpu().numpy().tolist() == [0, 1, 1, 1, 1, 1] and int(d_order[0]) == 0
            got_q, short = m.base_qual_counters()
            assert got_q[0].tolist() == [0, 0, 0, 17, case["template_bases"]] and short == 0, case["src"]
            assert int(got_q[1:].sum()) == 0 and m.unmatched_counts() == {}
            assert st.metrics().per_sample_array()[0].tolist() == [1, 1, 0]
            st.close()


def test_downsizing_while_another_thread_submits():
    """The table is cut down inside sgdx_sync of one thread while the other keeps launching batches on its own
    slot (post_mu orders them).  No drops, at least one cut, and the heavy keys keep their exact counts."""
    import threading
    rng = np.random.default_rng(17)
    w = sg.synth.Workload("dz", 16, 16, 8, 1, 2, n_reads=0, table_seed=3, read_seed=4)
    table = w.table()
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    heavy = acgt[rng.integers(0, 4, (8, 16))]
    st = _stage(w, table, unmatched_on_device=True, slots=2, most_unmatched_max_map_size=20_000,
                most_unmatched_downsize_to=64)
    batches = []
    for t in range(2):
        per_thread = []
        for b in range(12):
            noise = acgt[rng.integers(0, 4, (6_000, 16))]
            reps = np.repeat(heavy, 300, axis=0)
            x = np.concatenate([noise, reps])
            rng.shuffle(x)
            per_thread.append(np.ascontiguousarray(x))
        batches.append(per_thread)
    errors, undetermined = [], [0, 0]

    def work(slot):
        try:
            for x in batches[slot]:
                idx = st.demultiplex(x, slot=slot).sample_indices
                undetermined[slot] += int((idx == w.num_samples).sum())
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(s,)) for s in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    info = st.matcher.unmatched_info()
    assert info["downsizes"] >= 1 and info["dropped_reads"] == 0
    assert info["unmatched_reads"] == sum(undetermined)
    top = dict(st.matcher.unmatched_top(8))
    # every heavy key is unmatched for this table (random 16-mers against 16 samples) and was seen 2 * 12 * 300 times
    for h in heavy:
        assert top.get(h.tobytes()) == 2 * 12 * 300
    st.close()


def test_batcher_feeds_the_unmatched_table():
    """The host batching stage (1000-read chunks -> pinned super-batches -> sgdx_match_batch) with the unmatched
    table on: the host path counts the same keys as the oracle, chunk boundaries do not matter."""
    import ctypes as C
    from singular_demux_b200._lib import check, lib
    w = sg.synth.C1
    n = 200_000
    table = w.table()
    reads = w.reads_host(0, n, table)
    st = _stage(w, table, unmatched_on_device=True, slots=2)
    b = C.c_void_p()
    check(lib().sgdx_batcher_create(st.matcher.handle, 32_768, C.byref(b)))
    got = []

    def take(everything):
        while True:
            pi, pd, k = C.c_void_p(), C.c_void_p(), C.c_uint64(0)
            check(lib().sgdx_batcher_next(b, C.byref(pi), C.byref(pd), C.byref(k)))
            if k.value:
                got.append(np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint16)), (k.value,)).copy())
            if k.value == 0 or not everything:
                return

    for i, chunk in enumerate(np.array_split(reads, 200)):
        c = np.ascontiguousarray(chunk)
        check(lib().sgdx_batcher_push(b, c.ctypes.data, c.shape[0]))
        if i % 40 == 39:  # results are taken as the run goes (at most 2 x slots batches may wait)
            take(False)
    check(lib().sgdx_batcher_flush(b))
    take(True)
    lib().sgdx_batcher_destroy(b)
    idx = np.concatenate(got)
    want_idx = _oracle_idx(w, table, reads)
    assert np.array_equal(idx, want_idx)
    assert st.matcher.unmatched_counts() == orc.unmatched_counts(reads, want_idx, w.num_samples)
    rows = st.most_frequent_unmatched(5)
    assert len(rows) == 5 and all("+" in r.barcode for r in rows)
    st.close()
