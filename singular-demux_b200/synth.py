"""Seeded synthetic workloads through include/sgdx_synth.h.

The generator is counter-based and bit-identical on host and device, so tests regenerate on the CPU
exactly what the bench generated on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from ._lib import SgdxError, SynthParams, check, lib


def _frac(p: float) -> int:
    return min(int(round(p * 4294967296.0)), 0xFFFFFFFF)


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json config as concrete generator + matcher parameters."""
    name: str
    num_samples: int
    barcode_len: int
    index1_len: int            # 0 = single index
    max_mismatches: int
    min_delta: int
    free_ns: int = 1
    max_no_calls: Optional[int] = None
    n_reads: int = 1_000_000
    table_seed: int = 0xB200
    read_seed: int = 0x5EED
    p_true: float = 0.92
    p_hop: float = 0.03
    p_sub: float = 0.005
    p_n: float = 0.001
    p_x: float = 0.0
    zipf: bool = True
    min_dist: int = 3

    @property
    def index_lengths(self) -> Optional[Tuple[int, int]]:
        return (self.index1_len, self.barcode_len - self.index1_len) if self.index1_len else None

    def params(self) -> SynthParams:
        return SynthParams(self.read_seed, self.num_samples, self.barcode_len, self.index1_len, int(self.zipf),
                           _frac(self.p_true), _frac(self.p_hop if self.index1_len else 0.0), _frac(self.p_sub),
                           _frac(self.p_n), _frac(self.p_x))

    def table(self) -> np.ndarray:
        out = np.zeros((self.num_samples, self.barcode_len), dtype=np.uint8)
        rc = lib().sgdx_synth_table(self.table_seed, self.num_samples, self.barcode_len, self.index1_len,
                                    self.min_dist, out.ctypes.data)
        if rc != 0:
            raise SgdxError(f"could not draw {self.num_samples} barcodes of length {self.barcode_len} "
                            f"with min distance {self.min_dist}")
        return out

    def reads_host(self, first: int, n: int, table: Optional[np.ndarray] = None, threads: int = 8) -> np.ndarray:
        table = self.table() if table is None else np.ascontiguousarray(table, dtype=np.uint8)
        out = np.empty((n, self.barcode_len), dtype=np.uint8)
        p = self.params()
        check(lib().sgdx_synth_reads_host(C.byref(p), table.ctypes.data, first, n, out.ctypes.data, threads),
              "sgdx_synth_reads_host")
        return out

    def reads_device(self, first: int, n: int, d_table_ptr: int, d_out_ptr: int, stream: int = 0) -> None:
        p = self.params()
        check(lib().sgdx_synth_reads_device(C.byref(p), d_table_ptr, first, n, d_out_ptr, stream or None),
              "sgdx_synth_reads_device")


# BASELINE.json configs: C1..C5
C1 = Workload("C1 96x(8+8) -m1 -d2", 96, 16, 8, 1, 2, n_reads=1_000_000, table_seed=0xB201, read_seed=0x5EEE)
C2 = Workload("C2 384x(10+10) defaults", 384, 20, 10, 1, 2, n_reads=100_000_000, table_seed=0xB202, read_seed=0x5EEF)
C3 = Workload("C3 1536x(12+12) -m3 -d1", 1536, 24, 12, 3, 1, n_reads=500_000_000, table_seed=0xB203, read_seed=0x5EF0)
C4 = Workload("C4 1x8 inline 2% N", 1, 8, 0, 1, 2, n_reads=100_000_000, table_seed=0xB204, read_seed=0x5EF1,
              p_true=0.95, p_hop=0.0, p_n=0.02)
C5 = Workload("C5 384x(10+10) 2B reads / 8 GPUs", 384, 20, 10, 1, 2, n_reads=2_000_000_000, table_seed=0xB202,
              read_seed=0x5EEF)
CONFIGS = {"C1": C1, "C2": C2, "C3": C3, "C4": C4, "C5": C5}
