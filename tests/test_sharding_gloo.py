"""N > 1 host logic on CPU: two gloo ranks each own a shard of the reads, compute their shard's counters (here
with the CPU oracle standing in for the per-GPU kernel), and rank 0 sums them; the sum must equal the single-pass
counters (metrics.rs:315-328: merging is a plain sum, so any sharding is exact)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    import sgdx_loader
    sg = sgdx_loader.load()
    w = sg.synth.Workload("shard", 96, 16, 8, 1, 2, n_reads=30_001, p_n=0.01, p_hop=0.06)
    table = w.table()
    return sg, w, table, w.reads_host(0, w.n_reads, table, threads=2)


def _oracle_metrics(sg, w, table, reads):
    from oracle import oracle as orc
    cfg = orc.Config(w.num_samples, w.barcode_len, w.max_mismatches, w.min_delta, w.free_ns, -1,
                     orc.MATCHER_CACHED_HAMMING, w.index1_len, w.barcode_len - w.index1_len)
    res = orc.COracle(cfg, table).run(reads, threads=2)
    hop = sg.SampleBarcodeHopTracker(w.index1_len, w.barcode_len - w.index1_len, dict(res["hops"]))
    return sg.DemuxedGroupMetrics.from_device(res["per_sample"], res["total_templates"], hop)


def _oracle_idx(w, table, reads):
    from oracle import oracle as orc
    cfg = orc.Config(w.num_samples, w.barcode_len, w.max_mismatches, w.min_delta, w.free_ns, -1,
                     orc.MATCHER_CACHED_HAMMING, w.index1_len, w.barcode_len - w.index1_len)
    return orc.COracle(cfg, table).run(reads, threads=2)["idx"]


def _quals(n):
    return np.random.default_rng(3).integers(33, 75, (n, 40), dtype=np.uint8)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sg, w, table, reads = _case()
        first, last = sg.sharding.shard_range(reads.shape[0], rank, world)
        mine = _oracle_metrics(sg, w, table, reads[first:last])
        merged = sg.sharding.gather_metrics(mine, dst=0)
        assert (merged is None) == (rank != 0)
        # the stages either side of the assignment: unmatched maps and base-quality counters merge by sum as well
        from oracle import oracle as orc
        res_idx = _oracle_idx(w, table, reads[first:last])
        um = orc.unmatched_counts(reads[first:last], res_idx, w.num_samples)
        rows = sg.sharding.gather_unmatched(um, n=50, index1_length=w.index1_len, dst=0)
        quals = _quals(reads.shape[0])[first:last]
        bq, _ = orc.base_qual_counts(quals, None, [(0, 40, "T")], res_idx, w.num_samples)
        bq_all = sg.sharding.gather_base_qual(bq, dst=0)
        assert (rows is None) == (bq_all is None) == (rank != 0)
        if rank == 0:
            with open(out_path + ".unmatched", "w") as f:
                f.write(repr([(r.barcode, r.count) for r in rows]))
            np.save(out_path + ".bq.npy", bq_all)
        if rank == 0:
            np.save(out_path, merged.per_sample_array())
            with open(out_path + ".hops", "w") as f:
                f.write(repr(sorted(merged.sample_barcode_hop_tracker.count_tracker.items())))
            with open(out_path + ".total", "w") as f:
                f.write(str(merged.total_templates))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    import sgdx_loader
    sg = sgdx_loader.load()
    for n in (0, 1, 7, 8, 100_003):
        for world in (1, 2, 3, 8):
            cuts = [sg.sharding.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sg.sharding.shard_range(10, 2, 2)


def test_merge_metrics_is_a_plain_sum():
    sg, w, table, reads = _case()
    whole = _oracle_metrics(sg, w, table, reads)
    parts = [_oracle_metrics(sg, w, table, reads[a:b]) for a, b in
             (sg.sharding.shard_range(reads.shape[0], r, 5) for r in range(5))]
    merged = sg.sharding.merge_metrics(parts)
    assert np.array_equal(merged.per_sample_array(), whole.per_sample_array())
    assert merged.total_templates == whole.total_templates == reads.shape[0]
    assert merged.sample_barcode_hop_tracker.count_tracker == whole.sample_barcode_hop_tracker.count_tracker
    assert sg.sharding.gather_metrics(whole) is whole  # no process group: identity


def test_two_gloo_ranks_merge_to_single_pass(tmp_path):
    sg, w, table, reads = _case()
    whole = _oracle_metrics(sg, w, table, reads)
    out = str(tmp_path / "merged.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert np.array_equal(np.load(out), whole.per_sample_array())
    assert open(out + ".total").read() == str(reads.shape[0])
    assert open(out + ".hops").read() == repr(sorted(whole.sample_barcode_hop_tracker.count_tracker.items()))
    from oracle import oracle as orc
    idx = _oracle_idx(w, table, reads)
    want_rows = sg.sharding.merge_unmatched([orc.unmatched_counts(reads, idx, w.num_samples)], 50, w.index1_len)
    assert open(out + ".unmatched").read() == repr([(r.barcode, r.count) for r in want_rows])
    assert "+" in want_rows[0].barcode
    want_bq, _ = orc.base_qual_counts(_quals(reads.shape[0]), None, [(0, 40, "T")], idx, w.num_samples)
    assert np.array_equal(np.load(out + ".bq.npy"), want_bq)
