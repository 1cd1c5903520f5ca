"""Counters: host-side mirror of the hot-path part of /root/reference/src/lib/metrics.rs.

  DemuxedGroupSampleMetrics            metrics.rs:407-437
  DemuxedGroupMetrics.update_with      metrics.rs:315-328   (merge by summation: exact under any sharding)
  SampleBarcodeHopTracker              metrics.rs:617-677
  BarcodeCount / UnmatchedCounter      metrics.rs:131-206   (host-side merge of the device tables' exports)
  BaseQualCounter                      metrics.rs:468-497   (filled from sgdx_base_qual_counters)
The integer columns come from device counters (sgdx_counters / sgdx_hop_export); derived float columns
are out of scope (written by the host's unchanged TSV writer, metrics.rs:440-465).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence, Dict, List, Optional

import numpy as np


@dataclass
class DemuxedGroupSampleMetrics:
    perfect_matches: int = 0
    one_mismatch_matches: int = 0
    total_matches: int = 0

    def update_with(self, other: "DemuxedGroupSampleMetrics") -> None:
        self.perfect_matches += other.perfect_matches
        self.one_mismatch_matches += other.one_mismatch_matches
        self.total_matches += other.total_matches


@dataclass
class BaseQualCounter:
    """metrics.rs:468-481; update_with_self = metrics.rs:489-496.  The per-base update (metrics.rs:499-518)
    runs on the device (sgdx_qual_counts_device)."""
    q20_bases: int = 0
    q30_bases: int = 0
    index_bases_qual_sum: int = 0
    index_bases_total_seen: int = 0
    bases: int = 0

    def update_with_self(self, other: "BaseQualCounter") -> None:
        self.q20_bases += other.q20_bases
        self.q30_bases += other.q30_bases
        self.index_bases_qual_sum += other.index_bases_qual_sum
        self.index_bases_total_seen += other.index_bases_total_seen
        self.bases += other.bases

    @staticmethod
    def from_row(row) -> "BaseQualCounter":
        return BaseQualCounter(*[int(x) for x in row])


@dataclass
class SampleBarcodeHopTracker:
    """count_tracker: concatenated observed barcode -> count (metrics.rs:621-623)."""
    index1_length: int
    index2_length: int
    count_tracker: Dict[bytes, int] = field(default_factory=dict)

    def update_with_self(self, other: "SampleBarcodeHopTracker") -> None:  # metrics.rs:633-638
        for k, v in other.count_tracker.items():
            self.count_tracker[k] = self.count_tracker.get(k, 0) + v

    def __len__(self) -> int:
        return len(self.count_tracker)

    def is_empty(self) -> bool:
        return not self.count_tracker

    def barcode_counts(self) -> List["BarcodeCount"]:
        """Rows of sample_barcode_hop_metrics.tsv: `index1+index2`, descending count (metrics.rs:374-387)."""
        rows = [BarcodeCount((k[:self.index1_length] + b"+" + k[self.index1_length:]).decode(), v)
                for k, v in self.count_tracker.items()]
        rows.sort(key=lambda r: (-r.count, r.barcode))
        return rows


@dataclass(frozen=True)
class BarcodeCount:  # metrics.rs:194-206
    barcode: str
    count: int


@dataclass
class DemuxedGroupMetrics:
    per_sample_metrics: List[DemuxedGroupSampleMetrics]
    total_templates: int = 0
    num_reads_failing_filters: int = 0      # filters stay on the host (demux.rs:470-480)
    num_reads_filtered_as_control: int = 0
    sample_barcode_hop_tracker: Optional[SampleBarcodeHopTracker] = None

    @staticmethod
    def from_device(per_sample: np.ndarray, total_templates: int, hop: Optional[SampleBarcodeHopTracker]):
        ps = [DemuxedGroupSampleMetrics(int(r[1]), int(r[2]), int(r[0])) for r in per_sample]
        return DemuxedGroupMetrics(ps, int(total_templates), 0, 0, hop)

    def update_with(self, other: "DemuxedGroupMetrics") -> None:  # metrics.rs:315-328
        self.num_reads_failing_filters += other.num_reads_failing_filters
        self.num_reads_filtered_as_control += other.num_reads_filtered_as_control
        self.total_templates += other.total_templates
        for s, o in zip(self.per_sample_metrics, other.per_sample_metrics):
            s.update_with(o)
        if self.sample_barcode_hop_tracker is not None and other.sample_barcode_hop_tracker is not None:
            self.sample_barcode_hop_tracker.update_with_self(other.sample_barcode_hop_tracker)

    def per_sample_array(self) -> np.ndarray:
        return np.array([[m.total_matches, m.perfect_matches, m.one_mismatch_matches]
                         for m in self.per_sample_metrics], dtype=np.uint64)


def delimit_barcode(concatenated: bytes, segment_lengths: Optional[Sequence[int]]) -> bytes:
    """ExtractedBarcode::add_bases (demux.rs:56-62): the B segments of a template joined by `+`.  Whatever the
    lengths do not cover is the last segment (header-barcode mode has barcodes of any length)."""
    if not segment_lengths:
        return concatenated
    parts, pos = [], 0
    for ln in segment_lengths:
        if pos >= len(concatenated):
            break
        parts.append(concatenated[pos:pos + ln])
        pos += ln
    if pos < len(concatenated):
        parts.append(concatenated[pos:])
    return b"+".join(parts)


class UnmatchedCounter:
    """metrics.rs:131-190: counts delimited unmatched barcodes; when the map exceeds max_map_size it is
    cut down to the downsize_to most frequent keys.  Fed from the sample-index array."""

    def __init__(self, max_map_size: int = 5_000_000, downsize_to: int = 5_000):
        self.max_map_size, self.downsize_to = max_map_size, downsize_to
        self.counts: Dict[bytes, int] = {}

    def insert_counts(self, keys: np.ndarray, counts: np.ndarray) -> None:
        for i in range(keys.shape[0]):
            k = keys[i].tobytes()
            self.counts[k] = self.counts.get(k, 0) + int(counts[i])
            if len(self.counts) > self.max_map_size:
                self.downsize()

    def downsize(self) -> None:
        top = sorted(self.counts.items(), key=lambda kv: -kv[1])[:self.downsize_to]
        self.counts = dict(top)

    def most_frequent(self, n: int, index1_length: int = 0,
                      segment_lengths: Optional[Sequence[int]] = None) -> List[BarcodeCount]:
        """segment_lengths = lengths of the sample-barcode (B) segments in FASTQ order, then segment order: the
        reference writes the delimited barcode, `+` between every two of them (demux.rs:56-62, pushed at :553-555).
        index1_length is the two-segment shorthand (one B segment per index read)."""
        if segment_lengths is None and index1_length:
            segment_lengths = [index1_length]  # the rest of the key is the second segment
        rows = sorted(self.counts.items(), key=lambda kv: (-kv[1], kv[0]))[:n]
        return [BarcodeCount(delimit_barcode(k, segment_lengths).decode(errors="replace"), v) for k, v in rows]
