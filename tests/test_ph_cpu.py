"""Tables of the perfect-hash kernel (sgdx_match_ph, csrc/sgdx_ph.cu) against the oracle, without a GPU.

sgdx_ph_probe_host (include/sgdx_synth.h) runs the same host code sgdx_create runs to build the two-level perfect
hash of the match neighbourhood, then answers compact records with the kernel's fast-path arithmetic (hash ->
bucket multiplier -> untagged entry -> distance test).  For every read made of A/C/G/T only that answer must be
Matcher::find's (matcher.rs:273-303 / 121-262) -- including tables whose barcodes are so close that ties and the
delta rule reject reads near them, which is where the precomputed verdict (not the distance test) decides."""
import ctypes as C

import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def probe(table, M, D, matcher, reads):
    S, L = table.shape
    cfg = sg._lib.Config(S, L, M, D, 1, -1, matcher, 0, 0, 0, 1, 0)
    rec = sg.pack_compact_host(reads)
    idx = np.empty(len(reads), dtype=np.uint16)
    dist = np.empty(len(reads), dtype=np.uint8)
    info = (C.c_uint32 * 9)()
    tb = np.ascontiguousarray(table)
    rc = sg.lib().sgdx_ph_probe_host(C.byref(cfg), tb.ctypes.data, rec.ctypes.data, len(reads), idx.ctypes.data,
                                     dist.ctypes.data, info)
    assert rc == 0
    return idx, dist, list(info)


def reads_near(table, rng, n, max_subs):
    """Reads at 0..max_subs substitutions from random barcodes, plus uniform random reads."""
    S, L = table.shape
    reads = table[rng.integers(0, S, n)].copy()
    nsub = rng.integers(0, max_subs + 1, n)
    for r in range(n):
        pos = rng.choice(L, min(nsub[r], L), replace=False)
        reads[r, pos] = LETTERS[rng.integers(0, 4, len(pos))]
    reads[-n // 10:] = LETTERS[rng.integers(0, 4, (n // 10, L))]
    return reads


def oracle_run(table, M, D, kind, reads):
    S, L = table.shape
    cfg = orc.Config(S, L, M, D, 1, -1, kind, 0, 0)
    got = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=2, use_cache=False)
    return got["idx"], got["dist"]


@pytest.mark.parametrize("cfg_name,M,D", [("C1", 1, 2), ("C2", 1, 2), ("C2", 0, 5), ("C1", 0, 0), ("C4", 1, 2),
                                          ("C4", 2, 0)])
def test_baseline_tables_equal_the_oracle(cfg_name, M, D):
    w = sg.synth.CONFIGS[cfg_name]
    table = w.table()
    S, L = table.shape
    rng = np.random.default_rng(S * 31 + M * 7 + D)
    reads = reads_near(table, rng, 20000, M + 3)
    for kind, matcher in ((orc.MATCHER_CACHED_HAMMING, 1), (orc.MATCHER_PRECOMPUTE, 2)):
        idx, dist, info = probe(table, M, D, matcher, reads)
        assert info[0] == 1, info
        want_idx, want_dist = oracle_run(table, M, D, kind, reads)
        assert np.array_equal(idx, want_idx)
        assert np.array_equal(dist, want_dist)
    if cfg_name == "C2" and M == 1:
        assert info[1] == 384 * 61 and info[3] == 15  # 23 424 keys in 32 768 slots


@pytest.mark.parametrize("seed,S,L,M,D", [(1, 40, 6, 1, 1), (2, 40, 6, 2, 1), (3, 64, 8, 1, 2), (4, 64, 8, 2, 0), (5, 12, 5, 1, 0),
                                          (6, 30, 7, 3, 1), (7, 200, 10, 1, 3), (8, 16, 4, 4, 1), (9, 100, 12, 1, 1),
                                          (10, 3, 21, 1, 2), (11, 500, 16, 1, 1)])
def test_close_tables_where_ties_and_delta_decide(seed, S, L, M, D):
    """Random distinct barcodes with no distance guarantee: many keys are within M of two samples, or have a
    competitor within delta, so their precomputed verdict is NoMatch although the distance test alone would pass."""
    rng = np.random.default_rng(seed)
    table = np.unique(LETTERS[rng.integers(0, 4, (S * 2, L))], axis=0)
    rng.shuffle(table)
    table = np.ascontiguousarray(table[:S])
    S = len(table)
    reads = reads_near(table, rng, 12000, M + 2)
    saw_rejected_near = False
    for kind, matcher in ((orc.MATCHER_CACHED_HAMMING, 1), (orc.MATCHER_PRECOMPUTE, 2)):
        idx, dist, info = probe(table, M, D, matcher, reads)
        if not info[0]:
            pytest.skip("configuration has too many keys for a perfect hash")
        want_idx, want_dist = oracle_run(table, M, D, kind, reads)
        assert np.array_equal(idx, want_idx)
        assert np.array_equal(dist, want_dist)
        near = (reads[:, None, :] != table[None, :, :]).sum(-1).min(1) <= M
        saw_rejected_near |= bool((near & (want_idx == S)).any())
    if L <= 8:
        assert saw_rejected_near, "the case this test exists for did not occur"


def test_exhaustive_short_barcodes_both_rule_sets():
    """Every one of the 4^6 strings, 9 rule settings, both rule sets."""
    rng = np.random.default_rng(99)
    L = 6
    table = np.unique(LETTERS[rng.integers(0, 4, (24, L))], axis=0)[:20]
    allreads = LETTERS[(np.arange(4 ** L)[:, None] >> (2 * np.arange(L))) & 3]
    for M in (0, 1, 2):
        for D in (0, 1, 3):
            for kind, matcher in ((orc.MATCHER_CACHED_HAMMING, 1), (orc.MATCHER_PRECOMPUTE, 2)):
                idx, dist, info = probe(table, M, D, matcher, allreads)
                assert info[0] == 1
                want_idx, want_dist = oracle_run(table, M, D, kind, allreads)
                assert np.array_equal(idx, want_idx), (M, D, kind)
                assert np.array_equal(dist, want_dist), (M, D, kind)


def test_reads_with_invalid_positions_are_deferred_and_big_key_sets_shrink():
    w = sg.synth.CONFIGS["C2"]
    table = w.table()
    reads = table[:64].copy()
    reads[::2, 3] = ord("N")
    reads[1::4, 7] = ord("R")
    idx, dist, info = probe(table, 1, 2, 1, reads)
    bad = (reads == ord("N")).any(1) | (reads == ord("R")).any(1)
    assert (idx[bad] == 0xFFFF).all() and (dist[bad] == 0xFE).all()
    assert np.array_equal(idx[~bad], np.arange(64)[~bad])
    # 384 x 20 bases at two mismatches: 384 * 1771 strings are too many -- the key set falls back to radius 1
    _, _, info2 = probe(table, 2, 1, 1, reads[:1])
    assert info2[0] == 1 and info2[6] == 1 and info2[1] == 384 * 61 and info2[7] == 0
    assert info[4] == 1 and info[5] >= 6  # min pairwise distance 6 > 2*1 + 2 + 1: single-N shortcut valid


def probe_wide(table, M, D, matcher, reads):
    S, L = table.shape
    cfg = sg._lib.Config(S, L, M, D, 1, -1, matcher, 0, 0, 0, 1, 0)
    rec = sg.pack_host(reads)
    idx = np.empty(len(reads), dtype=np.uint16)
    dist = np.empty(len(reads), dtype=np.uint8)
    info = (C.c_uint32 * 9)()
    tb = np.ascontiguousarray(table)
    rc = sg.lib().sgdx_ph_probe_host_wide(C.byref(cfg), tb.ctypes.data, rec.ctypes.data, len(reads), idx.ctypes.data,
                                          dist.ctypes.data, info)
    assert rc == 0
    return idx, dist, list(info)


def _check_decided(idx, dist, info, table, M, reads, want_idx, want_dist):
    """Decided reads carry the oracle's verdict; the undecided ones (dist 0xFD) are exactly the plain reads further than
    the radius from every barcode; with radius == M nothing is undecided."""
    radius = max(info[6], info[7])  # what the two levels decide together
    undecided = (idx == 0xFFFF) & (dist == 0xFD)
    deferred = (idx == 0xFFFF) & (dist == 0xFE)
    decided = ~(undecided | deferred)
    assert np.array_equal(idx[decided], want_idx[decided]) and np.array_equal(dist[decided], want_dist[decided])
    plain = np.isin(reads, LETTERS).all(1)
    assert np.array_equal(deferred, ~plain)
    dmin = (reads[:, None, :] != table[None, :, :]).sum(-1).min(1)
    must = plain & (dmin > radius) & (radius < M)
    assert (undecided[must]).all() and (plain[undecided]).all()
    # a key whose precomputed verdict is NoMatch (tie, delta) is left to the general path too when the radius is short
    assert (want_idx[undecided & ~must] == len(table)).all() and (radius < M or not undecided.any())
    return radius


@pytest.mark.parametrize("cfg_name,M,D", [("C3", 3, 1), ("C3", 1, 2), ("C3", 0, 0), ("C2", 2, 1), ("C2", 3, 0), ("C1", 2, 1)])
def test_radius_smaller_than_max_mismatches_and_wide_keys(cfg_name, M, D):
    """Tables whose full neighbourhood does not fit: the key set shrinks to a smaller radius (C3: the 1536 barcodes
    themselves) and the rest is left to the general path; 24-base barcodes use the wide (16-byte record) keys."""
    w = sg.synth.CONFIGS[cfg_name]
    table = w.table()
    S, L = table.shape
    rng = np.random.default_rng(S + M)
    reads = reads_near(table, rng, 6000, M + 2)
    reads[::97, 3] = ord("N")
    for kind, matcher in ((orc.MATCHER_CACHED_HAMMING, 1), (orc.MATCHER_PRECOMPUTE, 2)):
        idx, dist, info = (probe_wide if L > 21 else probe)(table, M, D, matcher, reads)
        assert info[0] == 1, info
        want_idx, want_dist = oracle_run(table, M, D, kind, reads)
        radius = _check_decided(idx, dist, info, table, M, reads, want_idx, want_dist)
    if cfg_name == "C3":  # first level: the 1536 barcodes; second level (global memory): their 73 x 1536 neighbours
        assert info[6] == 0 and info[1] == 1536 and info[7] == min(M, 1) and (M == 0 or info[8] == 1536 * 73)
    if cfg_name == "C2" and M >= 2:  # 384 x 1771 strings within two substitutions fit neither level
        assert info[6] == 1 and info[1] == 384 * 61 and info[7] == 0


@pytest.mark.parametrize("seed,S,L,M,D", [(21, 50, 24, 1, 1), (22, 20, 32, 1, 2), (23, 8, 32, 2, 0), (24, 100, 22, 1, 3), (25, 300, 26, 1, 1)])
def test_wide_keys_on_close_tables(seed, S, L, M, D):
    """22..32-base barcodes without a distance guarantee (ties, delta rejections), all-T / all-G reads included: a read
    of 32 T's equals the dummy row's planes, which must not count as a match."""
    rng = np.random.default_rng(seed)
    table = np.unique(LETTERS[rng.integers(0, 4, (S * 2, L))], axis=0)
    rng.shuffle(table)
    table = np.ascontiguousarray(table[:S])
    # near-duplicates so that ties and competitors within delta exist
    for i in range(0, len(table) - 1, 5):
        table[i + 1] = table[i]
        table[i + 1, rng.integers(0, L)] = LETTERS[rng.integers(0, 4)]
    table = np.unique(table, axis=0)
    reads = reads_near(table, rng, 8000, M + 2)
    reads[0] = ord("T")
    reads[1] = ord("G")
    reads[2] = ord("A")
    for kind, matcher in ((orc.MATCHER_CACHED_HAMMING, 1), (orc.MATCHER_PRECOMPUTE, 2)):
        idx, dist, info = probe_wide(table, M, D, matcher, reads)
        assert info[0] == 1
        want_idx, want_dist = oracle_run(table, M, D, kind, reads)
        _check_decided(idx, dist, info, table, M, reads, want_idx, want_dist)


@pytest.mark.parametrize("cfg_name,M,D,kind", [("C3", 3, 1, orc.MATCHER_CACHED_HAMMING), ("C3", 2, 2, orc.MATCHER_CACHED_HAMMING),
                                               ("C2", 2, 1, orc.MATCHER_CACHED_HAMMING), ("C2", 2, 1, orc.MATCHER_PRECOMPUTE),
                                               ("C1", 3, 0, orc.MATCHER_PRECOMPUTE), ("C3", 3, 1, orc.MATCHER_PRECOMPUTE)])
def test_short_seed_index_rule_holds_in_the_oracle(cfg_name, M, D, kind):
    """The claim behind PhParams::ms_* (sgdx_match_phx): enumerate the samples that agree with the read on every valid
    position of one of max_mismatches + 1 chunks; with best = the closest of them (p valid-position mismatches, ninv
    invalid positions) the read is decided by that enumeration alone when there is no candidate, when p > M, or when
    2 p + min_delta + ninv < dmin -- the verdict of the rules applied to (best, second best of the enumeration) is then
    Matcher::find's.  Restated here in numpy and compared with the oracle, reads with no-calls included."""
    w = sg.synth.CONFIGS[cfg_name]
    table = w.table()
    S, L = table.shape
    pair = (table[:, None, :] != table[None, :, :]).sum(-1)
    np.fill_diagonal(pair, 99)
    dmin = int(pair.min())
    rng = np.random.default_rng(S + 10 * M + D + kind)
    n = 3000
    reads = reads_near(table, rng, n, M + 2)
    if w.index1_len:  # index hops: halves of two different samples
        a, b = rng.integers(0, S, 300), rng.integers(0, S, 300)
        reads[:300] = np.concatenate([table[a, :w.index1_len], table[b, w.index1_len:]], axis=1)
    nmask = rng.random((n, L)) < 0.02
    nmask[: n // 2] = False
    reads[nmask] = ord("N")
    valid = reads != ord("N")
    ninv = (~valid).sum(1)
    p_all = ((reads[:, None, :] != table[None, :, :]) & valid[:, None, :]).sum(-1)  # [n, S]
    C = M + 1
    bounds = [(c * L // C, (c + 1) * L // C) for c in range(C)]
    cand = np.zeros((n, S), dtype=bool)
    for a, b in bounds:
        cand |= (((reads[:, None, a:b] != table[None, :, a:b]) & valid[:, None, a:b]).sum(-1) == 0)
    assert (cand | (p_all > M)).all(), "pigeonhole: every sample within M shares a chunk"
    key = np.where(cand, (p_all + ninv[:, None]) * 65536 + np.arange(S)[None, :], 1 << 40)
    order = np.sort(key, axis=1)
    best, nxt = order[:, 0], (order[:, 1] if S > 1 else np.full(n, 1 << 40))
    none = best >= (1 << 40)
    bd, nd = best >> 16, np.where(nxt >= (1 << 40), 0xFFFF, nxt >> 16)
    pb = bd - ninv
    sure = none | (pb > M) | (2 * pb + D + ninv < dmin)
    assert sure.mean() > 0.5  # (reads at 0..M+2 substitutions, uniformly: far more borderline reads than real data)
    free_ns = 1
    if kind == orc.MATCHER_CACHED_HAMMING:
        rep = bd - np.minimum(ninv, free_ns)
        ok = ~none & (rep <= M) & (nd - bd > D)
    else:
        pc_limit = max(M, M + D - 1) if M + D else L
        nd_vis = np.where((ninv > 0) & (nd == M + 1), 0xFFFF, nd)  # (a second best at M + 1 is invisible: only matters unsure)
        rep = bd
        ok = ~none & (bd <= M) & ~((nd_vis <= bd + D) & (nd_vis <= pc_limit))
    want_idx = np.where(ok, best & 0xFFFF, S).astype(np.uint16)
    want_dist = np.where(ok, rep, 0xFF).astype(np.uint8)
    cfg = orc.Config(S, L, M, D, free_ns, -1, kind, 0, 0)
    got = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=2, use_cache=False)
    assert np.array_equal(got["idx"][sure], want_idx[sure])
    assert np.array_equal(got["dist"][sure], want_dist[sure])
