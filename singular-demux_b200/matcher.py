"""Matchers: host-side mirror of /root/reference/src/lib/matcher.rs over the CUDA C ABI.

Names, argument meaning and results follow the reference:
  MatchResult                      matcher.rs:27-47
  MatcherKind                      matcher.rs:49-72
  trait Matcher::find              matcher.rs:75-77
  CachedHammingDistanceMatcher     matcher.rs:80-109   (rule set: matcher.rs:273-348)
  PreComputeMatcher                matcher.rs:121-262
  select_matcher                   run.rs:196-198
The only addition is ``find_batch`` (the per-chunk/batch form a GPU needs).  All matching runs in the
sm_100a kernels; there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import BaseQual, Config, SgdxError, Totals, UnmatchedInfo, check, lib

UNDETERMINED_NAME = "Undetermined"  # matcher.rs:19


@dataclass(frozen=True)
class SampleMetadata:
    """The three fields of sample_metadata.rs::SampleMetadata the matcher reads."""
    sample_id: str
    barcode: bytes
    ordinal: int = 0


@dataclass(frozen=True)
class MatchResult:
    """matcher.rs:27-47: Match{sample_index, hamming_dist, barcode} | NoMatch{barcode}."""
    barcode: bytes
    sample_index: Optional[int] = None
    hamming_dist: Optional[int] = None

    @staticmethod
    def Match(sample_index: int, hamming_dist: int, barcode: bytes) -> "MatchResult":
        return MatchResult(bytes(barcode), int(sample_index), int(hamming_dist))

    @staticmethod
    def NoMatch(barcode: bytes) -> "MatchResult":
        return MatchResult(bytes(barcode))

    def is_match(self) -> bool:
        return self.sample_index is not None

    def is_no_match(self) -> bool:
        return not self.is_match()


class MatcherKind(enum.Enum):
    CachedHammingDistance = "cached-hamming-distance"
    PreCompute = "pre-compute"

    @staticmethod
    def from_str(s: str) -> "MatcherKind":  # matcher.rs:61-72 (clap ArgEnum names, case-insensitive)
        for v in MatcherKind:
            if v.value == s.lower():
                return v
        raise ValueError(f"Invalid variant: {s}")


def select_matcher(first_barcode_len: int, override: Optional[MatcherKind] = None) -> MatcherKind:
    """run.rs:196-198."""
    if first_barcode_len <= 12 or override == MatcherKind.PreCompute:
        return MatcherKind.PreCompute
    return MatcherKind.CachedHammingDistance


def _barcodes_of(samples) -> Tuple[bytes, int, int]:
    bcs = []
    for s in samples:
        b = s.barcode if isinstance(s, SampleMetadata) else s
        bcs.append(b.encode() if isinstance(b, str) else bytes(b))
    if not bcs:
        raise SgdxError("a GPU matcher needs at least one sample (the no-demux mode never reaches the matcher)")
    L = len(bcs[0])
    if any(len(b) != L for b in bcs):
        raise SgdxError("all sample barcodes must have the same length (run.rs:96-99)")
    return b"".join(bcs), len(bcs), L


def as_barcode_array(barcodes, L: int) -> np.ndarray:
    """[n, L] uint8 view of a batch given as ndarray / bytes / list of bytes."""
    if isinstance(barcodes, np.ndarray):
        a = np.ascontiguousarray(barcodes, dtype=np.uint8)
        return a.reshape(-1, L)
    if isinstance(barcodes, (bytes, bytearray)):
        return np.frombuffer(bytes(barcodes), dtype=np.uint8).reshape(-1, L)
    rows = [b.encode() if isinstance(b, str) else bytes(b) for b in barcodes]
    if any(len(r) != L for r in rows):
        raise SgdxError(f"device batches need barcodes of exactly {L} bytes")
    return np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), L)


def pack_compact_host(barcodes: np.ndarray, threads: int = 1, out: Optional[np.ndarray] = None) -> np.ndarray:
    """sgdx_pack_compact_host: (n, L) ASCII barcodes -> n compact 64-bit records (format: include/sgdx.h)."""
    arr = np.ascontiguousarray(barcodes, dtype=np.uint8)
    n, L = arr.shape
    rec = np.empty(n, dtype=np.uint64) if out is None else out
    check(lib().sgdx_pack_compact_host(arr.ctypes.data, n, L, rec.ctypes.data, threads), "sgdx_pack_compact_host")
    return rec


def dense_bytes(barcode_len: int) -> int:
    """SGDX_DENSE_BYTES: bytes of one dense record (2 bits per base)."""
    return (2 * barcode_len + 7) // 8


def pack_dense_host(barcodes: np.ndarray, threads: int = 1, exc_cap: Optional[int] = None):
    """sgdx_pack_dense_host: (n, L) ASCII barcodes, L <= 32 -> (dense bytes [n * dense_bytes(L)], exception read
    indices [k] uint32, their records: compact records [k] uint64 for L <= 21, sgdx_read records [k, 4] uint32 above).
    exc_cap: room for exceptions (default: every read)."""
    arr = np.ascontiguousarray(barcodes, dtype=np.uint8)
    n, L = arr.shape
    cap = n if exc_cap is None else int(exc_cap)
    dense = np.empty(n * dense_bytes(L), dtype=np.uint8)
    eidx = np.empty(max(cap, 1), dtype=np.uint32)
    erec = np.empty(max(cap, 1), dtype=np.uint64) if L <= 21 else np.empty((max(cap, 1), 4), dtype=np.uint32)
    k = C.c_uint64(0)
    check(lib().sgdx_pack_dense_host(arr.ctypes.data, n, L, dense.ctypes.data, eidx.ctypes.data, erec.ctypes.data, cap,
                                     C.byref(k), threads), "sgdx_pack_dense_host")
    return dense, eidx[:k.value], erec[:k.value]


def pack_host(barcodes: np.ndarray, threads: int = 1) -> np.ndarray:
    """sgdx_pack_host: (n, L) ASCII barcodes, L <= 32 -> (n, 4) uint32 sgdx_read records {lo, hi, nmask, xmask}."""
    arr = np.ascontiguousarray(barcodes, dtype=np.uint8)
    n, L = arr.shape
    rec = np.empty((n, 4), dtype=np.uint32)
    check(lib().sgdx_pack_host(arr.ctypes.data, n, L, rec.ctypes.data, threads), "sgdx_pack_host")
    return rec


class GpuMatcher:
    """A matcher bound to one CUDA device through sgdx_create (owns an sgdx handle)."""

    kind: MatcherKind = MatcherKind.CachedHammingDistance

    def __init__(self, samples: Sequence[Union[SampleMetadata, bytes, str]], max_mismatches: int, min_delta: int,
                 free_ns: int, *, max_no_calls: Optional[int] = None, index_lengths: Optional[Tuple[int, int]] = None,
                 device: int = 0, slots: int = 2, kernel: int = _lib.KERNEL_AUTO):
        table, S, L = _barcodes_of(samples)
        self.samples = list(samples)
        self.max_mismatches, self.min_delta, self.free_ns = int(max_mismatches), int(min_delta), int(free_ns)
        self.max_no_calls = max_no_calls
        self.index_lengths = index_lengths
        self.num_samples, self.barcode_len, self.device = S, L, device
        i1, i2 = index_lengths if index_lengths else (0, 0)
        cap = 0xFFFFFFFF
        self._cfg = Config(S, L, min(self.max_mismatches, cap), min(self.min_delta, cap), min(self.free_ns, cap),
                           -1 if max_no_calls is None else int(max_no_calls),
                           _lib.MATCHER_PRECOMPUTE if self.kind == MatcherKind.PreCompute else _lib.MATCHER_CACHED_HAMMING,
                           i1, i2, device, slots, kernel)
        self._h = C.c_void_p()
        self._table = table
        check(lib().sgdx_create(C.byref(self._cfg), table, C.byref(self._h)), "sgdx_create")
        eff = Config()
        check(lib().sgdx_get_config(self._h, C.byref(eff)))
        self.slots = eff.slots
        self.kernel = eff.kernel

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            lib().sgdx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    # -- trait Matcher ----------------------------------------------------------------------------
    def find(self, barcode: Union[bytes, str]) -> MatchResult:
        """matcher.rs:75-77. One barcode = one tiny device batch; use find_batch for throughput."""
        bc = barcode.encode() if isinstance(barcode, str) else bytes(barcode)
        if len(bc) != self.barcode_len:
            # only reachable with --sample-barcode-in-fastq-header; PreCompute keys all have length L
            # (NoMatch), and the scalar zip()-truncating Hamming case stays with the host
            if self.kind == MatcherKind.PreCompute:
                return MatchResult.NoMatch(bc)
            raise SgdxError(f"barcode of {len(bc)} bytes, matcher built for {self.barcode_len}")
        idx, dist = self.find_batch(np.frombuffer(bc, dtype=np.uint8).reshape(1, -1))
        if int(idx[0]) == self.num_samples:
            return MatchResult.NoMatch(bc)
        return MatchResult.Match(int(idx[0]), int(dist[0]), bc)

    def find_batch(self, barcodes, slot: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """Assign a batch: returns (sample index per read, S = no match; hamming_dist or NO_DIST)."""
        arr = as_barcode_array(barcodes, self.barcode_len)
        n = arr.shape[0]
        idx = np.empty(n, dtype=np.uint16)
        dist = np.empty(n, dtype=np.uint8)
        self.submit(arr, idx, dist, slot)
        self.sync(slot)
        return idx, dist

    # -- asynchronous / device-resident forms -----------------------------------------------------
    def submit(self, arr: np.ndarray, out_idx: np.ndarray, out_dist: Optional[np.ndarray], slot: int = 0) -> None:
        n = arr.shape[0] if arr.ndim == 2 else arr.size // self.barcode_len
        check(lib().sgdx_match_batch(self._h, slot, arr.ctypes.data, n, out_idx.ctypes.data,
                                     out_dist.ctypes.data if out_dist is not None else None), "sgdx_match_batch")

    def submit_ptr(self, in_ptr: int, n: int, idx_ptr: int, dist_ptr: int, slot: int = 0) -> None:
        check(lib().sgdx_match_batch(self._h, slot, in_ptr, n, idx_ptr, dist_ptr or None), "sgdx_match_batch")

    def sync(self, slot: int = 0) -> float:
        ms = C.c_float(0)
        check(lib().sgdx_sync(self._h, slot, C.byref(ms)), "sgdx_sync")
        return ms.value

    def match_device(self, d_ascii_ptr: int, n: int, d_idx_ptr: int, d_dist_ptr: int = 0, slot: int = 0) -> None:
        check(lib().sgdx_match_device(self._h, slot, d_ascii_ptr, n, d_idx_ptr, d_dist_ptr or None), "sgdx_match_device")

    def pack_device(self, d_ascii_ptr: int, n: int, d_packed_ptr: int, slot: int = 0) -> None:
        check(lib().sgdx_pack_device(self._h, slot, d_ascii_ptr, n, d_packed_ptr), "sgdx_pack_device")

    def match_packed_device(self, d_packed_ptr: int, n: int, d_idx_ptr: int, d_dist_ptr: int = 0, slot: int = 0) -> None:
        check(lib().sgdx_match_packed_device(self._h, slot, d_packed_ptr, n, d_idx_ptr, d_dist_ptr or None),
              "sgdx_match_packed_device")

    # -- compact records: one 64-bit word per read (barcode_len <= 21), include/sgdx.h ------------
    def find_batch_compact(self, barcodes, slot: int = 0, threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """find_batch through the compact path: host pack -> 8 bytes per read over PCIe -> sgdx_match_compact."""
        rec = pack_compact_host(as_barcode_array(barcodes, self.barcode_len), threads)
        idx = np.empty(rec.size, dtype=np.uint16)
        dist = np.empty(rec.size, dtype=np.uint8)
        self.submit_compact_ptr(rec.ctypes.data, rec.size, idx.ctypes.data, dist.ctypes.data, slot)
        self.sync(slot)
        return idx, dist

    def find_batch_dense(self, barcodes, slot: int = 0, threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """find_batch through dense records: 2 bits per base + an exception list over PCIe (barcode_len <= 32)."""
        dense, eidx, erec = pack_dense_host(as_barcode_array(barcodes, self.barcode_len), threads)
        n = dense.size // dense_bytes(self.barcode_len)
        idx = np.empty(n, dtype=np.uint16)
        dist = np.empty(n, dtype=np.uint8)
        self.submit_dense_ptr(dense.ctypes.data, n, eidx.ctypes.data, erec.ctypes.data, len(eidx), idx.ctypes.data,
                              dist.ctypes.data, slot)
        self.sync(slot)
        return idx, dist

    def submit_dense_ptr(self, dense_ptr: int, n: int, exc_idx_ptr: int, exc_rec_ptr: int, n_exc: int, idx_ptr: int,
                         dist_ptr: int, slot: int = 0) -> None:
        check(lib().sgdx_match_dense_batch(self._h, slot, dense_ptr, n, exc_idx_ptr or None, exc_rec_ptr or None, n_exc,
                                           idx_ptr, dist_ptr or None), "sgdx_match_dense_batch")

    def find_batch_packed(self, barcodes, slot: int = 0, threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """find_batch through 16-byte records packed on the host (any barcode_len): sgdx_match_packed_batch."""
        rec = pack_host(as_barcode_array(barcodes, self.barcode_len), threads)
        n = rec.shape[0]
        idx = np.empty(n, dtype=np.uint16)
        dist = np.empty(n, dtype=np.uint8)
        check(lib().sgdx_match_packed_batch(self._h, slot, rec.ctypes.data, n, idx.ctypes.data, dist.ctypes.data),
              "sgdx_match_packed_batch")
        self.sync(slot)
        return idx, dist

    def submit_packed_ptr(self, rec_ptr: int, n: int, idx_ptr: int, dist_ptr: int, slot: int = 0) -> None:
        check(lib().sgdx_match_packed_batch(self._h, slot, rec_ptr, n, idx_ptr, dist_ptr or None), "sgdx_match_packed_batch")

    def submit_compact_ptr(self, rec_ptr: int, n: int, idx_ptr: int, dist_ptr: int, slot: int = 0) -> None:
        check(lib().sgdx_match_compact_batch(self._h, slot, rec_ptr, n, idx_ptr, dist_ptr or None),
              "sgdx_match_compact_batch")

    def pack_compact_device(self, d_ascii_ptr: int, n: int, d_records_ptr: int, slot: int = 0) -> None:
        check(lib().sgdx_pack_compact_device(self._h, slot, d_ascii_ptr, n, d_records_ptr), "sgdx_pack_compact_device")

    def match_compact_device(self, d_records_ptr: int, n: int, d_idx_ptr: int, d_dist_ptr: int = 0, slot: int = 0) -> None:
        check(lib().sgdx_match_compact_device(self._h, slot, d_records_ptr, n, d_idx_ptr, d_dist_ptr or None),
              "sgdx_match_compact_device")

    def set_stream(self, slot: int, cuda_stream: int) -> None:
        check(lib().sgdx_set_stream(self._h, slot, cuda_stream or None), "sgdx_set_stream")

    def kernel_name(self) -> str:
        return lib().sgdx_kernel_name(self._h).decode()

    def compact_kernel_name(self) -> str:
        return lib().sgdx_compact_kernel_name(self._h).decode()

    def packed_kernel_name(self) -> str:
        return lib().sgdx_packed_kernel_name(self._h).decode()

    def path_flags(self) -> int:
        return int(lib().sgdx_path_flags(self._h))

    def launch_count(self) -> int:
        return int(lib().sgdx_launch_count(self._h))

    # -- counters ---------------------------------------------------------------------------------
    def counters(self) -> Tuple[np.ndarray, Dict[str, int]]:
        per = np.zeros((self.num_samples + 1, 3), dtype=np.uint64)
        t = Totals()
        check(lib().sgdx_counters(self._h, per.ctypes.data, C.byref(t)), "sgdx_counters")
        return per, {"total_templates": t.total_templates, "matched": t.matched, "undetermined": t.undetermined,
                     "hop_reads": t.hop_reads}

    def reset_counters(self) -> None:
        check(lib().sgdx_reset_counters(self._h), "sgdx_reset_counters")

    def hop_counts(self) -> Dict[bytes, int]:
        n = C.c_uint64(0)
        check(lib().sgdx_hop_count(self._h, C.byref(n)), "sgdx_hop_count")
        k = int(n.value)
        keys = np.zeros((max(k, 1), self.barcode_len), dtype=np.uint8)
        cnt = np.zeros(max(k, 1), dtype=np.uint64)
        w = C.c_uint64(0)
        check(lib().sgdx_hop_export(self._h, keys.ctypes.data, cnt.ctypes.data, max(k, 1), C.byref(w)), "sgdx_hop_export")
        return {keys[i].tobytes(): int(cnt[i]) for i in range(int(w.value))}


    # -- most-frequent unmatched barcodes (UnmatchedCounter, metrics.rs:131-190), counted on the device ----
    def unmatched_enable(self, max_map_size: int = 5_000_000, downsize_to: int = 5_000) -> None:
        check(lib().sgdx_unmatched_enable(self._h, max_map_size, downsize_to), "sgdx_unmatched_enable")

    def unmatched_info(self) -> Dict[str, int]:
        i = UnmatchedInfo()
        check(lib().sgdx_unmatched_info_get(self._h, C.byref(i)), "sgdx_unmatched_info_get")
        return {"distinct_keys": i.distinct_keys, "unmatched_reads": i.unmatched_reads,
                "dropped_reads": i.dropped_reads, "downsizes": i.downsizes}

    def _unmatched_pairs(self, fn, what: str, k: int):
        keys = np.zeros((max(k, 1), self.barcode_len), dtype=np.uint8)
        cnt = np.zeros(max(k, 1), dtype=np.uint64)
        w = C.c_uint64(0)
        check(fn(keys, cnt, w), what)
        return [(keys[i].tobytes(), int(cnt[i])) for i in range(int(w.value))]

    def unmatched_counts(self) -> Dict[bytes, int]:
        """Every concatenated unmatched barcode and its count (for merging shards)."""
        k = self.unmatched_info()["distinct_keys"]
        return dict(self._unmatched_pairs(
            lambda keys, cnt, w: lib().sgdx_unmatched_export(self._h, keys.ctypes.data, cnt.ctypes.data, max(k, 1),
                                                             C.byref(w)), "sgdx_unmatched_export", k))

    def unmatched_top(self, n: int):
        """The n most frequent (barcode, count) pairs, descending count (to_file, metrics.rs:177-190)."""
        return self._unmatched_pairs(
            lambda keys, cnt, w: lib().sgdx_unmatched_top(self._h, n, keys.ctypes.data, cnt.ctypes.data, C.byref(w)),
            "sgdx_unmatched_top", n)

    def unmatched_downsize(self) -> None:
        check(lib().sgdx_unmatched_downsize(self._h), "sgdx_unmatched_downsize")

    # -- segment extraction (demux.rs:360-417,504-522) ---------------------------------------------------------
    def _sources(self, d_base_ptrs: Sequence[int], strides: Sequence[int]):
        k = len(d_base_ptrs)
        ptrs = (C.c_void_p * k)(*[C.c_void_p(int(p)) for p in d_base_ptrs])
        st = (C.c_uint32 * k)(*[int(x) for x in strides])
        return k, ptrs, st

    def extract_device(self, d_base_ptrs: Sequence[int], strides: Sequence[int], structures, n: int,
                       d_out_ptr: int, slot: int = 0) -> None:
        from .read_structure import c_segments
        k, ptrs, st = self._sources(d_base_ptrs, strides)
        segs, ns = c_segments(structures)
        check(lib().sgdx_extract_device(self._h, slot, k, ptrs, st, segs, ns, n, d_out_ptr), "sgdx_extract_device")

    def match_segments_device(self, d_base_ptrs: Sequence[int], strides: Sequence[int], structures, n: int,
                              d_idx_ptr: int, d_dist_ptr: int = 0, slot: int = 0) -> None:
        from .read_structure import c_segments
        k, ptrs, st = self._sources(d_base_ptrs, strides)
        segs, ns = c_segments(structures)
        check(lib().sgdx_match_segments_device(self._h, slot, k, ptrs, st, segs, ns, n, d_idx_ptr,
                                               d_dist_ptr or None), "sgdx_match_segments_device")

    # -- base-quality counters (metrics.rs:499-518, demux.rs:582-583) -------------------------------------------
    def qual_counts_device(self, d_quals_ptr: int, n: int, stride: int, structure, d_idx_ptr: int,
                           d_lens_ptr: int = 0, slot: int = 0) -> None:
        from .read_structure import c_segments
        segs, ns = c_segments([structure])
        check(lib().sgdx_qual_counts_device(self._h, slot, d_quals_ptr, n, stride, d_lens_ptr or None, segs, ns,
                                            d_idx_ptr), "sgdx_qual_counts_device")

    def base_qual_counters(self) -> Tuple[np.ndarray, int]:
        """[(S+1), 5] = q20_bases, q30_bases, index_bases_qual_sum, index_bases_total_seen, bases; short reads."""
        arr = (BaseQual * (self.num_samples + 1))()
        short = C.c_uint64(0)
        check(lib().sgdx_base_qual_counters(self._h, arr, C.byref(short)), "sgdx_base_qual_counters")
        out = np.array([[a.q20_bases, a.q30_bases, a.index_bases_qual_sum, a.index_bases_total_seen, a.bases]
                        for a in arr], dtype=np.uint64)
        return out, int(short.value)

    # -- read routing (demux.rs:637-645: per_sample_reads) -------------------------------------------------------
    def route_device(self, d_idx_ptr: int, n: int, d_order_ptr: int, d_offsets_ptr: int, slot: int = 0) -> None:
        check(lib().sgdx_route_device(self._h, slot, d_idx_ptr, n, d_order_ptr, d_offsets_ptr), "sgdx_route_device")


    # -- Demultiplex::demultiplex for a device-resident batch (demux.rs:445-584) -----------------------------------
    def demultiplex_device(self, fastqs, structures, n: int, d_idx_ptr: int, d_dist_ptr: int = 0, d_order_ptr: int = 0,
                           d_offsets_ptr: int = 0, slot: int = 0) -> None:
        """fastqs: one (d_bases_ptr, d_quals_ptr or 0, d_lens_ptr or 0, stride) per FASTQ; structures: their
        ReadStructures.  Extraction + assignment (+ unmatched table) + quality counters (+ routing) on one stream."""
        from ._lib import DemuxOut, FastqView
        from .read_structure import c_segments
        views = (FastqView * len(fastqs))()
        for i, (b, q, ln, stride) in enumerate(fastqs):
            views[i] = FastqView(int(b), int(q) or None, int(ln) or None, int(stride))
        segs, ns = c_segments(structures)
        out = DemuxOut(int(d_idx_ptr), int(d_dist_ptr) or None, int(d_order_ptr) or None, int(d_offsets_ptr) or None)
        check(lib().sgdx_demultiplex_device(self._h, slot, views, len(fastqs), segs, ns, n, C.byref(out)),
              "sgdx_demultiplex_device")


class CachedHammingDistanceMatcher(GpuMatcher):
    """matcher.rs:80-109. The reference's per-thread LRU memo is an implementation detail of its CPU
    scan (semantically transparent); the kernel re-evaluates every read instead."""
    kind = MatcherKind.CachedHammingDistance

    def __init__(self, samples, max_mismatches: int, min_delta: int, free_ns: int, **kw):
        super().__init__(samples, max_mismatches, min_delta, free_ns, **kw)


class PreComputeMatcher(GpuMatcher):
    """matcher.rs:121-262. free_ns is accepted and ignored, as in the reference (matcher.rs:199)."""
    kind = MatcherKind.PreCompute

    def __init__(self, samples, max_mismatches: int, min_delta: int, free_ns: int, **kw):
        super().__init__(samples, max_mismatches, min_delta, free_ns, **kw)
