"""ctypes binding of libsgdx.so (include/sgdx.h, include/sgdx_synth.h).

The CUDA library is the product: if it is missing, or there is no CUDA device when a matcher is
created, this fails loudly.  There is no CPU fallback and nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SGDX_LIB") or os.path.join(_HERE, "libsgdx.so")  # SGDX_LIB: experiment builds (tools/)

MATCHER_AUTO, MATCHER_CACHED_HAMMING, MATCHER_PRECOMPUTE = 0, 1, 2
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_SEEDED = 0, 1, 2
NO_DIST = 0xFF
PATH_SEED_INDEX, PATH_EXACT, PATH_NEIGHBOURS, PATH_HOP_DIRECT = 1, 2, 4, 8
PATH_PERFECT_HASH = 64
PATH_SHORT_SEED = 128
COMPACT_MAX_LEN = 21
MAX_BARCODE_LEN = 64
MAX_PACKED_LEN = 32
MAX_SAMPLES = 8192


class SgdxError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("num_samples", C.c_uint32), ("barcode_len", C.c_uint32), ("max_mismatches", C.c_uint32),
                ("min_delta", C.c_uint32), ("free_ns", C.c_uint32), ("max_no_calls", C.c_int32),
                ("matcher", C.c_uint32), ("index1_len", C.c_uint32), ("index2_len", C.c_uint32),
                ("device", C.c_int32), ("slots", C.c_uint32), ("kernel", C.c_uint32)]


class Totals(C.Structure):
    _fields_ = [("total_templates", C.c_uint64), ("matched", C.c_uint64), ("undetermined", C.c_uint64),
                ("hop_reads", C.c_uint64)]


class UnmatchedInfo(C.Structure):
    _fields_ = [("distinct_keys", C.c_uint64), ("unmatched_reads", C.c_uint64), ("dropped_reads", C.c_uint64),
                ("downsizes", C.c_uint64)]


class Segment(C.Structure):
    """sgdx_segment: one read-structure segment of one FASTQ."""
    _fields_ = [("source", C.c_uint32), ("offset", C.c_uint32), ("length", C.c_uint32), ("kind", C.c_uint32)]


class BaseQual(C.Structure):
    _fields_ = [("q20_bases", C.c_uint64), ("q30_bases", C.c_uint64), ("index_bases_qual_sum", C.c_uint64),
                ("index_bases_total_seen", C.c_uint64), ("bases", C.c_uint64)]


class FastqView(C.Structure):
    _fields_ = [("d_bases", C.c_void_p), ("d_quals", C.c_void_p), ("d_lens", C.c_void_p), ("stride", C.c_uint32)]


class DemuxOut(C.Structure):
    _fields_ = [("d_sample_idx", C.c_void_p), ("d_hamming_dist", C.c_void_p), ("d_order", C.c_void_p),
                ("d_offsets", C.c_void_p)]


SEG_TO_END = 0xFFFFFFFF
MAX_SOURCES, MAX_SEGMENTS = 8, 32


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("num_samples", C.c_uint32), ("barcode_len", C.c_uint32),
                ("index1_len", C.c_uint32), ("zipf", C.c_uint32), ("p_true", C.c_uint32), ("p_hop", C.c_uint32),
                ("p_sub", C.c_uint32), ("p_n", C.c_uint32), ("p_x", C.c_uint32)]


# every symbol include/*.h declares; tests check each one is exported
SYMBOLS = [
    "sgdx_last_error", "sgdx_version", "sgdx_create", "sgdx_destroy", "sgdx_get_config", "sgdx_alloc_host",
    "sgdx_free_host", "sgdx_match_batch", "sgdx_sync", "sgdx_match_device", "sgdx_pack_device",
    "sgdx_match_packed_device", "sgdx_pack_host", "sgdx_match_packed_batch", "sgdx_pack_compact_host", "sgdx_pack_compact_device", "sgdx_match_compact_batch", "sgdx_pack_dense_host", "sgdx_match_dense_batch",
    "sgdx_match_compact_device", "sgdx_set_stream", "sgdx_kernel_name", "sgdx_compact_kernel_name", "sgdx_packed_kernel_name", "sgdx_ph_probe_host_wide", "sgdx_path_flags", "sgdx_launch_count", "sgdx_counters", "sgdx_reset_counters",
    "sgdx_hop_count", "sgdx_hop_export", "sgdx_batcher_create", "sgdx_batcher_destroy", "sgdx_batcher_push", "sgdx_batcher_push_n",
    "sgdx_batcher_flush", "sgdx_batcher_set_compact", "sgdx_batcher_next", "sgdx_measure_int_peak",
    "sgdx_unmatched_enable", "sgdx_unmatched_info_get", "sgdx_unmatched_export", "sgdx_unmatched_top",
    "sgdx_unmatched_downsize", "sgdx_extract_device", "sgdx_match_segments_device", "sgdx_qual_counts_device",
    "sgdx_base_qual_counters", "sgdx_route_device", "sgdx_demultiplex_device",
    "sgdx_synth_table", "sgdx_synth_reads_host", "sgdx_synth_reads_device", "sgdx_ph_probe_host",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SgdxError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). sgdx has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.sgdx_last_error.restype = C.c_char_p
    L.sgdx_version.restype = C.c_char_p
    L.sgdx_create.argtypes = [C.POINTER(Config), vp, C.POINTER(vp)]
    L.sgdx_destroy.argtypes = [vp]
    L.sgdx_destroy.restype = None
    L.sgdx_get_config.argtypes = [vp, C.POINTER(Config)]
    L.sgdx_alloc_host.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.sgdx_free_host.argtypes = [vp]
    L.sgdx_match_batch.argtypes = [vp, u32, vp, u64, vp, vp]
    L.sgdx_sync.argtypes = [vp, u32, C.POINTER(C.c_float)]
    L.sgdx_match_device.argtypes = [vp, u32, vp, u64, vp, vp]
    L.sgdx_pack_device.argtypes = [vp, u32, vp, u64, vp]
    L.sgdx_match_packed_device.argtypes = [vp, u32, vp, u64, vp, vp]
    L.sgdx_pack_host.argtypes = [vp, u64, u32, vp, u32]
    L.sgdx_match_packed_batch.argtypes = [vp, u32, vp, u64, vp, vp]
    L.sgdx_pack_compact_host.argtypes = [vp, u64, u32, vp, u32]
    L.sgdx_pack_compact_device.argtypes = [vp, u32, vp, u64, vp]
    L.sgdx_match_compact_batch.argtypes = [vp, u32, vp, u64, vp, vp]
    L.sgdx_pack_dense_host.argtypes = [vp, u64, u32, vp, vp, vp, u64, vp, u32]
    L.sgdx_match_dense_batch.argtypes = [vp, u32, vp, u64, vp, vp, u64, vp, vp]
    L.sgdx_match_compact_device.argtypes = [vp, u32, vp, u64, vp, vp]
    L.sgdx_set_stream.argtypes = [vp, u32, vp]
    L.sgdx_kernel_name.argtypes = [vp]
    L.sgdx_kernel_name.restype = C.c_char_p
    L.sgdx_compact_kernel_name.argtypes = [vp]
    L.sgdx_compact_kernel_name.restype = C.c_char_p
    L.sgdx_packed_kernel_name.argtypes = [vp]
    L.sgdx_packed_kernel_name.restype = C.c_char_p
    L.sgdx_ph_probe_host_wide.argtypes = [C.POINTER(Config), vp, vp, u64, vp, vp, vp]
    L.sgdx_path_flags.argtypes = [vp]
    L.sgdx_path_flags.restype = u32
    L.sgdx_launch_count.argtypes = [vp]
    L.sgdx_launch_count.restype = u64
    L.sgdx_counters.argtypes = [vp, vp, C.POINTER(Totals)]
    L.sgdx_reset_counters.argtypes = [vp]
    L.sgdx_hop_count.argtypes = [vp, C.POINTER(u64)]
    L.sgdx_hop_export.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    L.sgdx_batcher_create.argtypes = [vp, u64, C.POINTER(vp)]
    L.sgdx_batcher_destroy.argtypes = [vp]
    L.sgdx_batcher_destroy.restype = None
    L.sgdx_batcher_push.argtypes = [vp, vp, u64]
    L.sgdx_batcher_push_n.argtypes = [vp, vp, u64, vp]
    L.sgdx_batcher_flush.argtypes = [vp]
    L.sgdx_batcher_set_compact.argtypes = [vp, i32]
    L.sgdx_batcher_next.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u64)]
    L.sgdx_unmatched_enable.argtypes = [vp, u64, u64]
    L.sgdx_unmatched_info_get.argtypes = [vp, C.POINTER(UnmatchedInfo)]
    L.sgdx_unmatched_export.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    L.sgdx_unmatched_top.argtypes = [vp, u64, vp, vp, C.POINTER(u64)]
    L.sgdx_unmatched_downsize.argtypes = [vp]
    L.sgdx_extract_device.argtypes = [vp, u32, u32, C.POINTER(vp), C.POINTER(u32), C.POINTER(Segment), u32, u64, vp]
    L.sgdx_match_segments_device.argtypes = [vp, u32, u32, C.POINTER(vp), C.POINTER(u32), C.POINTER(Segment), u32,
                                             u64, vp, vp]
    L.sgdx_qual_counts_device.argtypes = [vp, u32, vp, u64, u32, vp, C.POINTER(Segment), u32, vp]
    L.sgdx_base_qual_counters.argtypes = [vp, C.POINTER(BaseQual), C.POINTER(u64)]
    L.sgdx_route_device.argtypes = [vp, u32, vp, u64, vp, vp]
    L.sgdx_demultiplex_device.argtypes = [vp, u32, C.POINTER(FastqView), u32, C.POINTER(Segment), u32, u64,
                                          C.POINTER(DemuxOut)]
    L.sgdx_measure_int_peak.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sgdx_synth_table.argtypes = [u64, u32, u32, u32, u32, vp]
    L.sgdx_synth_reads_host.argtypes = [C.POINTER(SynthParams), vp, u64, u64, vp, i32]
    L.sgdx_synth_reads_device.argtypes = [C.POINTER(SynthParams), vp, u64, u64, vp, vp]
    L.sgdx_ph_probe_host.argtypes = [C.POINTER(Config), vp, vp, u64, vp, vp, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("sgdx_last_error", "sgdx_version"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str = "sgdx") -> None:
    if rc != 0:
        msg = lib().sgdx_last_error()
        raise SgdxError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
