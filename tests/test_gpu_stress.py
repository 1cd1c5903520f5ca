"""Repeated launches of the queue-based kernels on N-heavy streams: every repetition must give the first repetition's
arrays and counters, and the first repetition the oracle's.  (A kernel whose queues race shows up as a handful of
differing reads once in a few dozen launches -- round 1's compact kernel was retired that way.  compute-sanitizer is
not available on the GPU pool.)"""
import dataclasses

import numpy as np
import pytest

import sgdx_loader
from oracle import oracle as orc

sg = sgdx_loader.load()
pytestmark = pytest.mark.gpu

REPS = 30


def _oracle(w, table, reads, M, D):
    il = w.index_lengths
    cfg = orc.Config(w.num_samples, w.barcode_len, M, D, 1, -1, orc.MATCHER_CACHED_HAMMING, il[0], il[1])
    return orc.COracle(cfg, [bytes(t) for t in table], force_closed_form=True).run(reads, threads=8, use_cache=True)


@pytest.mark.parametrize("cfg,M,D,form", [("C2", 1, 2, "compact"), ("C2", 2, 1, "compact"), ("C3", 3, 1, "packed"), ("C3", 1, 2, "packed")])
def test_repeated_launches_of_the_perfect_hash_kernels(cfg, M, D, form):
    import torch
    w = dataclasses.replace(sg.synth.CONFIGS[cfg], p_n=0.02, p_sub=0.03, p_hop=0.08, p_true=0.8)
    table = w.table()
    n = 400_003
    reads = w.reads_host(0, n, table)
    reads[::1000] = ord("N")
    want = _oracle(w, table, reads, M, D)
    m = sg.CachedHammingDistanceMatcher([bytes(t) for t in table], M, D, 1, index_lengths=w.index_lengths)
    if form == "compact":
        d_rec = torch.from_numpy(sg.pack_compact_host(reads).view(np.int64)).cuda()
        run = lambda i, d: m.match_compact_device(d_rec.data_ptr(), n, i.data_ptr(), d.data_ptr())
        assert m.compact_kernel_name() in ("sgdx_match_ph", "sgdx_match_phx")
    else:
        d_rec = torch.from_numpy(sg.pack_host(reads).view(np.int32)).cuda()
        run = lambda i, d: m.match_packed_device(d_rec.data_ptr(), n, i.data_ptr(), d.data_ptr())
        assert m.packed_kernel_name() == "sgdx_match_phx"
    for rep in range(REPS):
        d_idx = torch.full((n,), -1, dtype=torch.int16, device="cuda")
        d_dist = torch.full((n,), 0x77, dtype=torch.uint8, device="cuda")
        m.reset_counters()
        run(d_idx, d_dist)
        m.sync()
        idx, dist = d_idx.cpu().numpy().view(np.uint16), d_dist.cpu().numpy()
        bad = np.nonzero((idx != want["idx"]) | (dist != want["dist"]))[0]
        assert len(bad) == 0, (rep, len(bad), bad[:8], idx[bad[:8]], want["idx"][bad[:8]])
        per, totals = m.counters()
        assert np.array_equal(per, want["per_sample"]) and totals["total_templates"] == n, rep
        assert m.hop_counts() == want["hops"], rep
    m.close()


def test_repeated_launches_of_the_routing_kernels():
    import torch
    rng = np.random.default_rng(11)
    for S in (384, 1536, 3):
        n = 1_000_003
        idx = (rng.random(n) ** 2 * (S + 1)).astype(np.uint16)
        idx[rng.random(n) < 0.1] = S
        m = sg.CachedHammingDistanceMatcher([bytes(b) for b in sg.synth.Workload("r", S, 12, 0, 1, 1, table_seed=1, min_dist=1).table()], 1, 1, 1)
        d_idx = torch.from_numpy(idx.view(np.int16)).cuda()
        want_order, want_off = orc.route_reads(idx, S)
        for rep in range(REPS):
            d_order = torch.full((n,), -1, dtype=torch.int32, device="cuda")
            d_off = torch.zeros(S + 2, dtype=torch.int64, device="cuda")
            m.route_device(d_idx.data_ptr(), n, d_order.data_ptr(), d_off.data_ptr())
            m.sync(0)
            assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), want_off), (S, rep)
            assert np.array_equal(d_order.cpu().numpy().view(np.uint32), want_order), (S, rep)
        m.close()
