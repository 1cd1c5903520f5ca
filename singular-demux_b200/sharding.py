"""Multi-GPU sharding of the assignment path: reads shard by contiguous range, one process per GPU, no
data-path collective; per-GPU counters are additive and are summed on the host exactly like the reference's
fold/reduce of per-chunk metrics (/root/reference/src/lib/run.rs:275-306, metrics.rs:315-328,633-638).

``torch.distributed`` is used only as plumbing (gather of a few KB of counters to rank 0); it works the same
over ``gloo`` on CPU (tests) and ``nccl`` on GPUs (bench.py)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .metrics import DemuxedGroupMetrics, SampleBarcodeHopTracker


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[first, last) of rank's contiguous shard of n reads; shards differ by at most one read."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    q, r = divmod(n, world)
    first = rank * q + min(rank, r)
    return first, first + q + (1 if rank < r else 0)


def merge_metrics(parts) -> Optional[DemuxedGroupMetrics]:
    """Sum the metrics of several shards (any order)."""
    merged = None
    for m in parts:
        if m is None:
            continue
        if merged is None:
            hop = None
            if m.sample_barcode_hop_tracker is not None:
                t = m.sample_barcode_hop_tracker
                hop = SampleBarcodeHopTracker(t.index1_length, t.index2_length, dict(t.count_tracker))
            merged = DemuxedGroupMetrics.from_device(m.per_sample_array(), m.total_templates, hop)
            merged.num_reads_failing_filters = m.num_reads_failing_filters
            merged.num_reads_filtered_as_control = m.num_reads_filtered_as_control
        else:
            merged.update_with(m)
    return merged


def gather_metrics(metrics: DemuxedGroupMetrics, dst: int = 0, group=None) -> Optional[DemuxedGroupMetrics]:
    """Collect every rank's counters on rank ``dst`` and sum them there; other ranks get None.
    Without an initialised process group this is the identity (single GPU)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return metrics
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = {
        "per_sample": metrics.per_sample_array(),
        "total_templates": metrics.total_templates,
        "failing": metrics.num_reads_failing_filters,
        "control": metrics.num_reads_filtered_as_control,
        "hop": None if metrics.sample_barcode_hop_tracker is None else
        (metrics.sample_barcode_hop_tracker.index1_length, metrics.sample_barcode_hop_tracker.index2_length,
         metrics.sample_barcode_hop_tracker.count_tracker),
    }
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for g in gathered:
        hop = None if g["hop"] is None else SampleBarcodeHopTracker(g["hop"][0], g["hop"][1], dict(g["hop"][2]))
        m = DemuxedGroupMetrics.from_device(np.asarray(g["per_sample"], dtype=np.uint64), g["total_templates"], hop)
        m.num_reads_failing_filters, m.num_reads_filtered_as_control = g["failing"], g["control"]
        parts.append(m)
    return merge_metrics(parts)


def merge_unmatched(parts, n: int = 1000, index1_length: int = 0, max_map_size: int = 5_000_000,
                    downsize_to: int = 5_000, segment_lengths=None):
    """Sum the unmatched-barcode maps of several shards (each one GpuMatcher.unmatched_counts()) and return the
    rows of most_frequent_unmatched.tsv (metrics.rs:177-190).  The reference has one counter for the whole run
    (metrics.rs:79-100); per-GPU tables are additive, so below max_map_size distinct keys the merge is exact."""
    from .metrics import UnmatchedCounter
    c = UnmatchedCounter(max_map_size, downsize_to)
    for p in parts:
        for k, v in p.items():
            c.counts[k] = c.counts.get(k, 0) + int(v)
        if len(c.counts) > max_map_size:
            c.downsize()
    return c.most_frequent(n, index1_length, segment_lengths)


def _gather_to(obj, dst: int, group):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(obj, gathered, dst=dst, group=group)
    return gathered


def gather_unmatched(counts: Dict[bytes, int], n: int = 1000, index1_length: int = 0, dst: int = 0, group=None,
                     segment_lengths=None):
    """Every rank's unmatched map to rank ``dst`` -> merged top-n rows there, None elsewhere."""
    parts = _gather_to(counts, dst, group)
    return None if parts is None else merge_unmatched(parts, n, index1_length, segment_lengths=segment_lengths)


def gather_base_qual(counters: np.ndarray, dst: int = 0, group=None) -> Optional[np.ndarray]:
    """Sum of every rank's [(S+1), 5] base-quality counters (BaseQualCounter::update_with_self, metrics.rs:489-496)
    on rank ``dst``; None elsewhere."""
    parts = _gather_to(np.asarray(counters, dtype=np.uint64), dst, group)
    return None if parts is None else np.sum(np.stack(parts), axis=0, dtype=np.uint64)
