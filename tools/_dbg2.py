import sys, os, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import sgdx_loader
sg = sgdx_loader.load()
import test_gpu_compact as t
mode = sys.argv[1]; reps = int(sys.argv[2])
S,L,i1 = 97,20,10
rng = np.random.default_rng(S * 977 + L)
w, table, reads = t._random_case(rng, S, L, i1)
il = w.index_lengths
kind = sg.MatcherKind.CachedHammingDistance
M,D,F,mnc = 2,3,1,None
want = t._oracle(table, kind, M, D, F, mnc, il, reads)
nbad = 0
for rep in range(reps):
    st = t._stage(table, kind, M, D, F, mnc, il)
    got = st.demultiplex(reads, compact=(mode != "ascii"), pack_threads=2) if mode != "ascii" else st.demultiplex(reads)
    bad = np.nonzero((got.sample_indices != want["idx"]) | (got.hamming_dists != want["dist"]))[0]
    if len(bad):
        nbad += 1
        print(mode, "rep", rep, "flags", st.matcher.path_flags(), "mismatches", len(bad), [(int(b), int(got.sample_indices[b])) for b in bad[:6]], flush=True)
    st.close()
print(mode, os.environ.get("SGDX_NO_SHORT_SEED"), "failures", nbad, "of", reps)
