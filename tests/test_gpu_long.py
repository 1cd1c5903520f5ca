"""GPU parity of sgdx_match_long (csrc/sgdx_long.cu): barcodes of 33..64 bases -- the reference has no length cap
(matcher.rs:335-348 zips slices of any length) -- against the CPU oracle through the C ABI: index, distance, per-sample
counters, total_templates and the hop map, for both rule sets; plus the entry points that stop at 32 bases."""
import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
pytestmark = pytest.mark.gpu

ORC_KIND = {sg.MatcherKind.CachedHammingDistance: orc.MATCHER_CACHED_HAMMING, sg.MatcherKind.PreCompute: orc.MATCHER_PRECOMPUTE}
CLS = {sg.MatcherKind.CachedHammingDistance: sg.CachedHammingDistanceMatcher, sg.MatcherKind.PreCompute: sg.PreComputeMatcher}
LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def _case(rng, S, L, i1, n, close=False):
    table = np.unique(LETTERS[rng.integers(0, 4, (S * 2, L))], axis=0)
    rng.shuffle(table)
    table = np.ascontiguousarray(table[:S])
    if close:  # pairs one substitution apart: ties and delta rejections
        for i in range(0, len(table) - 1, 4):
            table[i + 1] = table[i]
            table[i + 1, rng.integers(0, L)] = LETTERS[rng.integers(0, 4)]
        table = np.ascontiguousarray(np.unique(table, axis=0))
    S = len(table)
    reads = table[rng.integers(0, S, n)].copy()
    sub = rng.random(reads.shape) < 0.03
    reads[sub] = LETTERS[rng.integers(0, 4, int(sub.sum()))]
    k = n // 10
    reads[:k] = LETTERS[rng.integers(0, 4, (k, L))]                      # random barcodes
    if i1:                                                               # index hops, both orientations where the lengths allow
        a, b = rng.integers(0, S, k), rng.integers(0, S, k)
        reads[k:2 * k] = np.concatenate([table[a, :i1], table[b, i1:]], axis=1)
        if 2 * i1 == L:
            reads[2 * k:2 * k + k // 2] = np.concatenate([table[b[:k // 2], i1:], table[a[:k // 2], :i1]], axis=1)
    reads[rng.random(reads.shape) < 0.01] = ord("N")
    reads[rng.random(reads.shape) < 0.0005] = ord("R")
    reads[-3] = ord("N")
    reads[-2, :i1 or 5] = ord("N")
    reads[-1] = ord("T")
    return table, np.ascontiguousarray(reads)


@pytest.mark.parametrize("S,L,i1", [(1, 33, 0), (40, 33, 16), (96, 40, 20), (384, 48, 24), (200, 64, 32), (60, 64, 0),
                                    (30, 50, 32), (1500, 36, 18), (7000, 34, 0)])
def test_long_barcodes_match_oracle(S, L, i1):
    rng = np.random.default_rng(S * 131 + L)
    n = 30_011 if S < 2000 else 6_007
    table, reads = _case(rng, S, L, i1, n)
    S = len(table)
    il = (i1, L - i1) if i1 else None
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        for (M, D, F, mnc) in [(1, 2, 1, None), (3, 1, 2, 2), (0, 0, 0, None), (5, 3, 1, None)]:
            m = CLS[kind]([bytes(t) for t in table], M, D, F, max_no_calls=mnc, index_lengths=il)
            assert m.kernel_name() == "sgdx_match_long" and m.packed_kernel_name() == "" and m.compact_kernel_name() == ""
            idx, dist = m.find_batch(reads)
            cfg = orc.Config(S, L, M, D, F, -1 if mnc is None else mnc, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
            want = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)
            assert np.array_equal(idx, want["idx"]) and np.array_equal(dist, want["dist"])
            per, totals = m.counters()
            assert np.array_equal(per, want["per_sample"]) and totals["total_templates"] == n
            if il:
                assert m.hop_counts() == want["hops"]
            m.close()


@pytest.mark.parametrize("L,i1", [(35, 0), (44, 22)])
def test_long_barcodes_on_close_tables_unaligned_and_ragged(L, i1):
    """Ties and delta rejections between barcodes one substitution apart; lengths that are not a multiple of four
    (byte loads), device-resident input at an odd address, no distance output, batches of 0 / 1 / 257 reads."""
    import torch
    rng = np.random.default_rng(L)
    table, reads = _case(rng, 64, L, i1, 20_001, close=True)
    S = len(table)
    il = (i1, L - i1) if i1 else None
    for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
        m = CLS[kind]([bytes(t) for t in table], 2, 1, 1, index_lengths=il)
        cfg = orc.Config(S, L, 2, 1, 1, -1, ORC_KIND[kind], il[0] if il else 0, il[1] if il else 0)
        o = orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True)
        total = 0
        for n in (0, 1, 257, len(reads)):
            part = reads[:n]
            buf = torch.zeros(n * L + 8, dtype=torch.uint8, device="cuda")
            buf[1:1 + n * L] = torch.from_numpy(part.reshape(-1)).cuda()
            d_idx = torch.full((n + 1,), -1, dtype=torch.int16, device="cuda")
            m.match_device(buf.data_ptr() + 1, n, d_idx.data_ptr())
            m.sync()
            want = o.run(part, threads=4, use_cache=False)
            assert np.array_equal(d_idx.cpu().numpy().view(np.uint16)[:n], want["idx"])
            total += n
        per, totals = m.counters()
        assert totals["total_templates"] == total
        m.close()


def test_what_stops_at_32_bases():
    table = LETTERS[np.random.default_rng(3).integers(0, 4, (8, 40))]
    m = sg.CachedHammingDistanceMatcher([bytes(t) for t in table], 1, 2, 1)
    with pytest.raises(sg.SgdxError):
        m.find_batch_packed(table)           # sgdx_read records hold 32 bases
    with pytest.raises(sg.SgdxError):
        m.unmatched_enable()                 # 96-bit keys
    m.close()
    with pytest.raises(sg.SgdxError):        # an index half of more than 32 bases
        sg.CachedHammingDistanceMatcher([bytes(t) for t in table], 1, 2, 1, index_lengths=(36, 4))
    with pytest.raises(sg.SgdxError):
        sg.CachedHammingDistanceMatcher([b"A" * 65], 1, 2, 1)
