"""sgdx -- B200-native per-read sample-barcode assignment for sgdemux-style demultiplexing.

Host-side mirror (Python) of the reference's interface for this one path, over the C ABI in
include/sgdx.h.  Import name: ``singular_demux_b200`` (the directory is ``singular-demux_b200``; use
``sgdx_loader.load()`` from the repo root, or put this directory's parent on sys.path under that name).
"""
from ._lib import (KERNEL_AUTO, KERNEL_SEEDED, KERNEL_DIRECT, LIB_PATH, MATCHER_AUTO, MATCHER_CACHED_HAMMING,
                   MATCHER_PRECOMPUTE, NO_DIST, SgdxError, lib)
from .matcher import (UNDETERMINED_NAME, CachedHammingDistanceMatcher, GpuMatcher, MatcherKind, MatchResult,
                      PreComputeMatcher, SampleMetadata, dense_bytes, pack_compact_host, pack_dense_host, pack_host,
                      select_matcher)
from .metrics import (BarcodeCount, BaseQualCounter, DemuxedGroupMetrics, DemuxedGroupSampleMetrics,
                      SampleBarcodeHopTracker, UnmatchedCounter, delimit_barcode)
from .read_structure import ReadSegment, ReadStructure, barcode_segment_lengths
from .demux import BarcodeAssignmentStage, DemuxedGroup, from_delimited
from . import read_structure, sharding, synth

__all__ = [n for n in dir() if not n.startswith("_")]
