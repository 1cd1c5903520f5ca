"""The assignment + counting step of Demultiplexer::demultiplex, batched for the GPU.

Mirrors /root/reference/src/lib/demux.rs:524-556,581 (N gate -> Matcher::find -> sample index ->
update_with_match -> hop check -> unmatched collection -> total_templates) and Demultiplexer::new's
wiring of matcher + hop checker (demux.rs:223-270).  Segment extraction, filters, header rewriting and
read routing stay with the host pipeline; this stage consumes the concatenated
sample barcodes they produce and returns the sample-index array they route by.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .matcher import GpuMatcher, SampleMetadata, as_barcode_array
from .metrics import DemuxedGroupMetrics, SampleBarcodeHopTracker, UnmatchedCounter


@dataclass
class DemuxedGroup:
    """What the host needs back per batch (demux.rs:637-645, minus the routed reads themselves)."""
    sample_indices: np.ndarray           # uint16, undetermined_sample_index for NoMatch
    hamming_dists: np.ndarray            # uint8, 0xFF for NoMatch
    unmatched: Optional[np.ndarray]      # [k, L] concatenated barcodes of the NoMatch reads (collect_unmatched)


def from_delimited(delimited: bytes) -> bytes:
    """ExtractedBarcode::from_delimited (demux.rs:42-53), the --sample-barcode-in-fastq-header path: the concatenated
    barcode the matcher sees is the header's barcode without its FIRST `+` (later ones stay)."""
    i = delimited.find(b"+")
    return delimited if i < 0 else delimited[:i] + delimited[i + 1:]


class BarcodeAssignmentStage:
    """samples: ALL samples, the last one must be Undetermined (demux.rs:221, 259); the matcher sees
    samples[:-1] (run.rs:201,220).  index_lengths = (index1_length, index2_length) when exactly two
    reads carry sample-barcode segments (metrics.rs:587-595), else None: no hop checker."""

    def __init__(self, samples: Sequence[SampleMetadata], matcher_cls, max_mismatches: int, min_delta: int,
                 free_ns: int, max_no_calls: Optional[int] = None, index_lengths: Optional[Tuple[int, int]] = None,
                 collect_unmatched: bool = True, device: int = 0, slots: int = 2, kernel: int = 0,
                 unmatched_on_device: bool = False, most_unmatched_max_map_size: int = 5_000_000,
                 most_unmatched_downsize_to: int = 5_000, barcode_segment_lengths: Optional[Sequence[int]] = None):
        if len(samples) < 2:
            raise ValueError("need at least one sample plus Undetermined")
        self.samples = list(samples)
        self.undetermined_sample_index = len(samples) - 1
        self.index_lengths = index_lengths
        # lengths of the B segments (FASTQ order, then segment order): where the delimited barcode gets its `+`
        # (demux.rs:56-62).  Default: one segment per barcode read, i.e. index_lengths, or a single segment.
        if barcode_segment_lengths is None:
            barcode_segment_lengths = list(index_lengths) if index_lengths else [len(samples[0].barcode)]
        if sum(barcode_segment_lengths) != len(samples[0].barcode):
            raise ValueError("barcode_segment_lengths must add up to the barcode length (run.rs:109-119)")
        self.barcode_segment_lengths = list(barcode_segment_lengths)
        self.collect_unmatched = collect_unmatched and not unmatched_on_device
        self.unmatched_on_device = unmatched_on_device
        self.matcher: GpuMatcher = matcher_cls(self.samples[:-1], max_mismatches, min_delta, free_ns,
                                               max_no_calls=max_no_calls, index_lengths=index_lengths,
                                               device=device, slots=slots, kernel=kernel)
        if unmatched_on_device:  # run.rs:185-190: UnmatchedCounterThread::new(channel, max_map_size, downsize_to)
            self.matcher.unmatched_enable(most_unmatched_max_map_size, most_unmatched_downsize_to)

    def most_frequent_unmatched(self, n: int = 1000) -> List["BarcodeCount"]:
        """Rows of most_frequent_unmatched.tsv (metrics.rs:177-190) from the device table; keys in their
        delimited form (demux.rs:56-62, 504-522: every sample-barcode segment joined to the next by `+`)."""
        from .metrics import BarcodeCount, delimit_barcode
        return [BarcodeCount(delimit_barcode(k, self.barcode_segment_lengths).decode(errors="replace"), c)
                for k, c in self.matcher.unmatched_top(n)]

    def get_undetermined_index(self) -> int:
        return self.undetermined_sample_index

    def demultiplex(self, barcodes, slot: int = 0, compact: bool = False, pack_threads: int = 1) -> DemuxedGroup:
        """compact=True: the batch crosses PCIe as one 64-bit record per read (barcode_len <= 21), packed here
        by sgdx_pack_compact_host; same results."""
        arr = as_barcode_array(barcodes, self.matcher.barcode_len)
        idx, dist = self.matcher.find_batch_compact(arr, slot, pack_threads) if compact else self.matcher.find_batch(arr, slot)
        unmatched = arr[idx == self.undetermined_sample_index] if self.collect_unmatched else None
        return DemuxedGroup(idx, dist, unmatched)

    def metrics(self) -> DemuxedGroupMetrics:
        """Cumulative counters of every batch this stage has processed (device-resident; read on demand)."""
        per, totals = self.matcher.counters()
        hop = None
        if self.index_lengths:
            hop = SampleBarcodeHopTracker(self.index_lengths[0], self.index_lengths[1], self.matcher.hop_counts())
        return DemuxedGroupMetrics.from_device(per, totals["total_templates"], hop)

    def close(self) -> None:
        self.matcher.close()


def count_unmatched(unmatched: np.ndarray, counter: Optional[UnmatchedCounter] = None) -> UnmatchedCounter:
    """Feed one batch's NoMatch barcodes to the top-N counter (the reference sends them to a counter
    thread, demux.rs:553-555 -> metrics.rs:79-100)."""
    counter = counter or UnmatchedCounter()
    if unmatched is not None and unmatched.shape[0]:
        u, c = np.unique(unmatched, axis=0, return_counts=True)
        counter.insert_counts(u, c)
    return counter
