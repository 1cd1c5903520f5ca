#!/usr/bin/env python3
"""Debug aid (GPU box): sweep the rule parameters of tests/test_gpu_parity.py::PARAMS for one table shape and
print, per setting, how many assignments / counter cells differ from the oracle.
  python tools/gpu_sweep.py S L index1_len [trials] [kernel]"""
import sys
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import sgdx_loader
sg = sgdx_loader.load()
import test_gpu_parity as t

S, L, i1 = (int(x) for x in sys.argv[1:4])
trials = int(sys.argv[4]) if len(sys.argv) > 4 else 1
kernel = int(sys.argv[5]) if len(sys.argv) > 5 else 2
rng = np.random.default_rng(S * 1000 + L)
w, table, reads = t._random_case(rng, S, L, i1)
n = reads.shape[0]
for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
    for (M, D, F, mnc) in t.PARAMS:
        want = t._oracle(table, kind, M, D, F, mnc, w.index_lengths, reads, force_closed_form=True)
        for _ in range(trials):
            st = t._stage(table, kind, M, D, F, mnc, w.index_lengths, kernel)
            idx = np.full(n, 0xEEEE, dtype=np.uint16)
            dist = np.full(n, 0xEE, dtype=np.uint8)
            st.matcher.submit(reads, idx, dist, 0)
            st.matcher.sync(0)
            bad = np.nonzero((idx != want["idx"]) | (dist != want["dist"]))[0]
            per, tot = st.matcher.counters()
            cbad = int((per != want["per_sample"]).sum())
            print(kind.value, (M, D, F, mnc), st.matcher.kernel_name(), "mismatches", len(bad), "counter cells", cbad,
                  "total", tot["total_templates"], "first bad", bad[:8].tolist(), flush=True)
            for b in bad[:3]:
                print("    read", b, reads[b].tobytes().decode(errors="replace"), "got", idx[b], dist[b], "want",
                      want["idx"][b], want["dist"][b])
            st.close()
