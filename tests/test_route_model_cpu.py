"""A plain-Python model of the whole-sector scatter that `sgdx_route_device` uses for large batches
(`sgdx_route_scatter_staged`, csrc/sgdx_post.cu): the rules by which a read id goes to the ring of
its bin, straight to the output, or out with a completed sector -- CTA ranges, warps taking turns, the sector phase
`a` of the output pointer, first sectors shared between two ranges, leftovers at the end of a range.  The model must
write every word of order[0..n) exactly once, nothing outside, and produce the stable partition of the read ids
(DemuxedGroup::per_sample_reads, demux.rs:578,637-645).  It states the algorithm for readers and guards its case
analysis without a GPU; the kernel itself is checked on the device
(tests/test_gpu_post.py::test_routing_staged_form_on_small_batches)."""
import numpy as np
import pytest


def staged_scatter_model(idx, bins, num_ranges, warps, a):
    """idx: bin per read (already clamped to bins - 1); `a`: words between the 32-byte boundary below order[] and
    order[0].  Returns order[] as the kernel would leave it."""
    n = len(idx)
    chunk = ((n + num_ranges - 1) // num_ranges + 31) // 32 * 32
    ranges = (n + chunk - 1) // chunk
    # histogram matrix [range][bin], exclusive scans: what sgdx_route_hist_cta and the two scan kernels leave
    matrix = np.stack([np.bincount(idx[g * chunk:min(n, (g + 1) * chunk)], minlength=bins) for g in range(ranges)])
    totals = matrix.sum(0)
    bin_offset = np.concatenate([[0], np.cumsum(totals)[:-1]])
    range_offset = np.cumsum(matrix, 0) - matrix
    order_a = np.full(n + 16, -1, dtype=np.int64)  # indexed by P = position + a
    writes = np.zeros(n + 16, dtype=np.int64)

    def store(P, read_id):
        order_a[P] = read_id
        writes[P] += 1

    for g in range(ranges):  # one CTA
        cur = range_offset[g] + bin_offset + a
        start = cur.copy()
        ring = np.full((bins, 8), -7, dtype=np.int64)
        lo, hi = g * chunk, min(n, (g + 1) * chunk)
        for tile in range(lo, hi, 32 * warps):
            for w in range(warps):  # the warps take their turns in this order
                reads = [tile + 32 * w + lane for lane in range(32)]
                lanes = []
                for lane, r in enumerate(reads):
                    if r >= hi:
                        lanes.append(None)
                        continue
                    b = idx[r]
                    peers = [k for k in range(32) if reads[k] < hi and idx[reads[k]] == b]
                    rank, cnt, base = peers.index(lane), len(peers), cur[b]
                    P, end = base + rank, base + cnt
                    s8 = P & ~7
                    complete = s8 + 8 <= end          # the step fills my sector
                    begun = s8 < base                  # an earlier step (or the range before) began it
                    lanes.append(dict(b=b, r=r, rank=rank, P=P, s8=s8, complete=complete, begun=begun,
                                      to_ring=(not complete) or begun, first=s8 == (base & ~7), end=end))
                for x in lanes:                        # the bin's lowest lane moves the cursor
                    if x and x["rank"] == 0:
                        cur[x["b"]] = x["end"]
                for x in lanes:                        # phase A: the step's first sector
                    if x and x["to_ring"] and x["first"]:
                        ring[x["b"], x["P"] & 7] = x["r"]
                for x in lanes:                        # a begun sector that the step completes goes out
                    if x and x["rank"] == 0 and x["begun"] and x["complete"]:
                        b, s8 = x["b"], x["s8"]
                        first_word = max(s8, start[b])  # below start[b]: the range before
                        for j in range(first_word, s8 + 8):
                            store(j, ring[b, j & 7])
                for x in lanes:                        # phase B: the sector the step leaves unfinished
                    if x and x["to_ring"] and not x["first"]:
                        ring[x["b"], x["P"] & 7] = x["r"]
                for x in lanes:                        # sectors filled from their first word: straight out
                    if x and not x["to_ring"]:
                        store(x["P"], x["r"])
        for b in range(bins):                          # leftovers of the range
            for j in range(max(cur[b] & ~7, start[b]), cur[b]):
                store(j, ring[b, j & 7])
    assert (writes[a:a + n] == 1).all(), "every word of order[] is written exactly once"
    assert writes[:a].sum() == 0 and writes[a + n:].sum() == 0, "nothing outside order[]"
    return order_a[a:a + n]


@pytest.mark.parametrize("seed", range(6))
def test_whole_sector_scatter_model_is_a_stable_partition(seed):
    rng = np.random.default_rng(seed)
    for _ in range(8):
        bins = int(rng.choice([1, 2, 3, 5, 17, 40]))
        n = int(rng.integers(1, 2500))
        if rng.random() < 0.5:
            idx = rng.integers(0, bins, n)
        else:  # one bin dominates: steps that fill several sectors at once
            idx = np.where(rng.random(n) < 0.85, 0, rng.integers(0, bins, n))
        got = staged_scatter_model(idx, bins, num_ranges=int(rng.integers(1, 5)), warps=int(rng.choice([2, 8])),
                                   a=int(rng.integers(0, 8)))
        assert np.array_equal(got, np.argsort(idx, kind="stable"))
