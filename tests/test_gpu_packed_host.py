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
