import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import sgdx_loader
sg = sgdx_loader.load()
import test_gpu_compact as t
S,L,i1 = 97,20,10
rng = np.random.default_rng(S * 977 + L)
w, table, reads = t._random_case(rng, S, L, i1)
il = w.index_lengths
wants = {}
nbad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
  for kind in (sg.MatcherKind.CachedHammingDistance, sg.MatcherKind.PreCompute):
    for (M, D, F, mnc) in ([(2, 3, 1, None)] if len(sys.argv) > 2 else t.PARAMS):
        key = (kind, M, D, F, mnc)
        if key not in wants:
            wants[key] = t._oracle(table, kind, M, D, F, mnc, il, reads)
        want = wants[key]
        st = t._stage(table, kind, M, D, F, mnc, il)
        got = st.demultiplex(reads, compact=True, pack_threads=2)
        bad = np.nonzero((got.sample_indices != want["idx"]) | (got.hamming_dists != want["dist"]))[0]
        m = st.metrics()
        cb = int((m.per_sample_array()!=want["per_sample"]).sum())
        hb = m.sample_barcode_hop_tracker.count_tracker != want["hops"]
        if len(bad) or cb or hb:
            nbad += 1
            print("rep", rep, kind.value, (M,D,F,mnc), "flags", st.matcher.path_flags(), "mismatches", len(bad), "counter cells", cb, "hops differ", hb, flush=True)
            for b in bad[:8]:
                print("   read", b, reads[b].tobytes(), "got", got.sample_indices[b], got.hamming_dists[b], "want", want["idx"][b], want["dist"][b])
        st.close()
print("stress done, failures:", nbad)
