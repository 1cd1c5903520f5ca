"""CPU oracle for sgdemux's per-read sample-barcode assignment path -- TEST INFRASTRUCTURE ONLY.

Two layers, both restating /root/reference/src/lib (citations are file:line in that tree):

* ``Literal*`` / ``literal_*``: pure-Python, line-by-line-in-spirit restatements (small cases only).
  ``literal_build_map`` rebuilds the PreComputeMatcher's map exactly the way matcher.rs:155-261 does
  (three passes over generated neighbour strings), so it is the ground truth for the closed form.
* ``COracle``: ctypes binding of ``sgdx_oracle.c`` (fast; multi-threaded; also the timed CPU baseline).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
module.  The product package never does.

Parity status: PINNED by tests/golden/reference_known_answers.json (every known-answer test the
reference holds for this path); the reference itself (Rust) cannot be built in this image.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MATCHER_CACHED_HAMMING = 0
MATCHER_PRECOMPUTE = 1
NO_DIST = 0xFF
ALLOWED_BASES = b"ACTGN"  # matcher.rs:16


# ------------------------------------------------------------------------------------------------
# literal pure-Python restatement
# ------------------------------------------------------------------------------------------------
def hamming_distance(alpha: bytes, beta: bytes, free_ns: int) -> int:
    """matcher.rs:335-348."""
    seen = 0
    dist = 0
    for a, b in zip(alpha, beta):
        no_call = a == 0x4E or b == 0x4E
        if no_call and seen < free_ns:
            seen += 1
        elif a != b or no_call:
            dist += 1
    return dist


def should_reject_delta(delta: int, min_delta: int) -> bool:
    """matcher.rs:266-270."""
    return delta <= min_delta


USIZE_MAX = (1 << 64) - 1


def find_independently(barcode: bytes, samples: Sequence[bytes], max_mismatches: int, min_delta: int,
                       free_ns: int) -> Optional[Tuple[int, int]]:
    """matcher.rs:273-303. Returns (sample_index, hamming_dist) or None for NoMatch."""
    best_sample = len(samples)
    best = USIZE_MAX
    nxt = USIZE_MAX
    for i, s in enumerate(samples):
        d = hamming_distance(barcode, s, free_ns)
        if d < best:
            best_sample, nxt, best = i, best, d
        elif d < nxt:
            nxt = d
    if best_sample == len(samples) or best > max_mismatches or should_reject_delta(nxt - best, min_delta):
        return None
    return best_sample, best


def _all_permutations(barcode: bytes, min_mm: int, max_mm: int):
    """matcher.rs:194-237 (free_ns is hard-wired to 0 inside, line 199).

    Yields the de-duplicated set of (string, distance) pairs."""
    k = min(max_mm, len(barcode))  # line 201
    stage1 = set()
    for locations in itertools.combinations(range(len(barcode)), k):
        choices = [bytes([c]) for c in barcode]
        for loc in locations:
            choices[loc] = ALLOWED_BASES
        for combo in itertools.product(*choices):
            bc = bytes(combo)
            d1 = hamming_distance(barcode, bc, 1)  # line 210: one free N in the filter
            if d1 >= min_mm:
                stage1.add((bc, d1))
    out = set()
    for bc, _ in stage1:
        # combinations(0) == one empty selection: the string itself, re-measured with free_ns = 0
        out.add((bc, hamming_distance(barcode, bc, 0)))
    return out


def literal_build_map(samples: Sequence[bytes], max_mismatches: int, min_delta: int,
                      ordinals: Optional[Sequence[int]] = None) -> Dict[bytes, Tuple[int, int]]:
    """PreComputeMatcher::build_map, matcher.rs:155-191 + pick_matching_sample_barcode 240-261.

    Returns {observed barcode: (sample ordinal, hamming_dist)}."""
    if ordinals is None:
        ordinals = list(range(len(samples)))
    table: Dict[bytes, set] = {}
    for s, o in zip(samples, ordinals):  # pass 1 (163-168)
        for read, d in _all_permutations(s, 0, max_mismatches):
            table.setdefault(read, set()).add((o, d))
    hi = (max_mismatches + min_delta - 1) & USIZE_MAX  # usize arithmetic wraps in release (175)
    for s, o in zip(samples, ordinals):  # pass 2 (171-181)
        for read, d in _all_permutations(s, max_mismatches + 1, hi):
            if read in table:
                table[read].add((o, d))
    result = {}
    for read, matches in table.items():  # pass 3 (184-190)
        best = min(d for _, d in matches)
        hits = [m for m in matches if m[1] == best]
        if len(hits) == 1:
            if sum(1 for _, d in matches if should_reject_delta(d - best, min_delta)) == 1:
                result[read] = hits[0]
    return result


def select_matcher(first_barcode_len: int, override: Optional[int] = None) -> int:
    """run.rs:196-198."""
    if first_barcode_len <= 12 or override == MATCHER_PRECOMPUTE:
        return MATCHER_PRECOMPUTE
    return MATCHER_CACHED_HAMMING


@dataclass
class Config:
    num_samples: int
    barcode_len: int
    max_mismatches: int = 1   # opts.rs:118-119
    min_delta: int = 2        # opts.rs:123-124
    free_ns: int = 1          # opts.rs:127-128
    max_no_calls: int = -1    # None (opts.rs:135-136)
    matcher: int = MATCHER_CACHED_HAMMING
    index1_len: int = 0
    index2_len: int = 0


class LiteralDemux:
    """demux.rs:524-556,581 + metrics.rs:428-437,575-614,658-676 in plain Python (small inputs)."""

    def __init__(self, cfg: Config, samples: Sequence[bytes]):
        assert len(samples) == cfg.num_samples
        self.cfg = cfg
        self.samples = [bytes(s) for s in samples]
        self.map = None
        if cfg.matcher == MATCHER_PRECOMPUTE:
            self.map = literal_build_map(self.samples, cfg.max_mismatches, cfg.min_delta)
        self.hop = None
        if cfg.index1_len > 0 and cfg.index2_len > 0:
            full = self.samples + [b"N" * cfg.barcode_len]  # Undetermined, sample_metadata.rs:378-383
            self.hop = ({b[:cfg.index1_len] for b in full}, {b[cfg.index1_len:] for b in full})

    def find(self, barcode: bytes) -> Optional[Tuple[int, int]]:
        if self.map is not None:
            return self.map.get(bytes(barcode))  # matcher.rs:129-140
        return find_independently(barcode, self.samples, self.cfg.max_mismatches, self.cfg.min_delta,
                                  self.cfg.free_ns)

    def run(self, reads: Sequence[bytes]):
        c = self.cfg
        S = c.num_samples
        per_sample = np.zeros((S + 1, 3), dtype=np.uint64)
        hops: Dict[bytes, int] = {}
        unmatched: Dict[bytes, int] = {}
        idx_out, dist_out = [], []
        total = 0
        for bc in reads:
            bc = bytes(bc)
            n = bc.count(b"N")  # demux.rs:78-80
            res = None if (c.max_no_calls >= 0 and n > c.max_no_calls) else self.find(bc)
            b = res[0] if res else S
            per_sample[b, 0] += 1
            if res and res[1] == 0:
                per_sample[b, 1] += 1
            elif res and res[1] == 1:
                per_sample[b, 2] += 1
            if res is None:
                if self.hop is not None and len(bc) == c.index1_len + c.index2_len:
                    i1, i2 = self.hop
                    if (bc[:c.index1_len] in i1 and bc[c.index1_len:] in i2) or \
                            (bc[c.index2_len:] in i1 and bc[:c.index2_len] in i2):
                        hops[bc] = hops.get(bc, 0) + 1
                unmatched[bc] = unmatched.get(bc, 0) + 1
            total += 1
            idx_out.append(b)
            dist_out.append(res[1] if res else NO_DIST)
        return {
            "idx": np.array(idx_out, dtype=np.uint16), "dist": np.array(dist_out, dtype=np.uint8),
            "per_sample": per_sample, "total_templates": total, "hops": hops, "unmatched": unmatched,
        }


# ------------------------------------------------------------------------------------------------
# ctypes binding of sgdx_oracle.c
# ------------------------------------------------------------------------------------------------
class _CConfig(C.Structure):
    _fields_ = [("num_samples", C.c_uint32), ("barcode_len", C.c_uint32), ("max_mismatches", C.c_uint32),
                ("min_delta", C.c_uint32), ("free_ns", C.c_uint32), ("max_no_calls", C.c_int32),
                ("matcher", C.c_uint32), ("index1_len", C.c_uint32), ("index2_len", C.c_uint32)]


def build(native: bool = False, force: bool = False) -> str:
    """Compile sgdx_oracle.c -> oracle/liboracle_sgdx[_native].so (gcc). Returns the path."""
    name = "liboracle_sgdx_native.so" if native else "liboracle_sgdx.so"
    out = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "sgdx_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        march = "native" if native else "x86-64-v3"
        cmd = ["gcc", "-O3", f"-march={march}", "-std=c11", "-fPIC", "-pthread", "-shared", "-o", out, src,
               "-lpthread"]
        subprocess.run(cmd, check=True, cwd=_HERE)
    return out


_LIBS: Dict[bool, C.CDLL] = {}


def lib(native: bool = False) -> C.CDLL:
    if native in _LIBS:
        return _LIBS[native]
    L = C.CDLL(build(native))
    u8p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
    L.orc_hamming_distance.restype = C.c_size_t
    L.orc_hamming_distance.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t]
    L.orc_should_reject_delta.restype = C.c_int
    L.orc_should_reject_delta.argtypes = [C.c_size_t, C.c_size_t]
    L.orc_select_matcher.restype = C.c_uint32
    L.orc_select_matcher.argtypes = [C.c_size_t, C.c_int]
    L.orc_find_independently.restype = C.c_int
    L.orc_find_independently.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t,
                                         C.c_size_t, C.c_size_t, C.c_size_t, szp, szp]
    L.orc_find_precompute_closed_form.restype = C.c_int
    L.orc_find_precompute_closed_form.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t,
                                                  C.c_size_t, C.c_size_t, szp, szp]
    L.orc_matcher_new.restype = C.c_void_p
    L.orc_matcher_new.argtypes = [C.POINTER(_CConfig), C.c_char_p, C.c_int]
    L.orc_matcher_free.argtypes = [C.c_void_p]
    L.orc_matcher_map_len.restype = C.c_size_t
    L.orc_matcher_map_len.argtypes = [C.c_void_p]
    L.orc_matcher_has_map.restype = C.c_int
    L.orc_matcher_has_map.argtypes = [C.c_void_p]
    L.orc_matcher_map_export.restype = C.c_size_t
    L.orc_matcher_map_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_matcher_find.restype = C.c_int
    L.orc_matcher_find.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, szp, szp]
    L.orc_hopmap_new.restype = C.c_void_p
    L.orc_hopmap_new.argtypes = [C.c_size_t]
    L.orc_hopmap_free.argtypes = [C.c_void_p]
    L.orc_hopmap_len.restype = C.c_size_t
    L.orc_hopmap_len.argtypes = [C.c_void_p]
    L.orc_hopmap_export.restype = C.c_size_t
    L.orc_hopmap_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_demux_batch.restype = C.c_int
    L.orc_demux_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_uint64), C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    _LIBS[native] = L
    return L


def _table_bytes(samples) -> bytes:
    if isinstance(samples, np.ndarray):
        return np.ascontiguousarray(samples, dtype=np.uint8).tobytes()
    return b"".join(bytes(s) for s in samples)


class COracle:
    """Matcher + demux/counting stage backed by sgdx_oracle.c."""

    def __init__(self, cfg: Config, samples, force_closed_form: bool = False, native: bool = False):
        self.cfg = cfg
        self.L = lib(native)
        self._table = _table_bytes(samples)
        assert len(self._table) == cfg.num_samples * cfg.barcode_len
        self._ccfg = _CConfig(cfg.num_samples, cfg.barcode_len, cfg.max_mismatches, cfg.min_delta, cfg.free_ns,
                              cfg.max_no_calls, cfg.matcher, cfg.index1_len, cfg.index2_len)
        self._m = self.L.orc_matcher_new(C.byref(self._ccfg), self._table, int(force_closed_form))

    def __del__(self):
        if getattr(self, "_m", None):
            self.L.orc_matcher_free(self._m)
            self._m = None

    @property
    def has_map(self) -> bool:
        return bool(self.L.orc_matcher_has_map(self._m))

    def map(self) -> Dict[bytes, Tuple[int, int]]:
        n = self.L.orc_matcher_map_len(self._m)
        Lb = self.cfg.barcode_len
        keys = np.zeros((max(n, 1), Lb), dtype=np.uint8)
        s = np.zeros(max(n, 1), dtype=np.uint32)
        d = np.zeros(max(n, 1), dtype=np.uint32)
        n2 = self.L.orc_matcher_map_export(self._m, keys.ctypes.data, s.ctypes.data, d.ctypes.data)
        assert n2 == n
        return {keys[i].tobytes(): (int(s[i]), int(d[i])) for i in range(n)}

    def find(self, barcode: bytes) -> Optional[Tuple[int, int]]:
        i, d = C.c_size_t(0), C.c_size_t(0)
        ok = self.L.orc_matcher_find(self._m, bytes(barcode), len(barcode), C.byref(i), C.byref(d))
        return (i.value, d.value) if ok else None

    def run(self, reads: np.ndarray, threads: int = 1, use_cache: bool = False, want_assignments: bool = True):
        """reads: uint8 array [n, L] of ASCII barcodes. Returns dict like LiteralDemux.run (+ 'seconds')."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n = reads.shape[0] if reads.ndim == 2 else reads.size // max(self.cfg.barcode_len, 1)
        S, Lb = self.cfg.num_samples, self.cfg.barcode_len
        idx = np.zeros(n, dtype=np.uint16) if want_assignments else None
        dist = np.zeros(n, dtype=np.uint8) if want_assignments else None
        per_sample = np.zeros((S + 1, 3), dtype=np.uint64)
        total = C.c_uint64(0)
        secs = C.c_double(0)
        hop_on = self.cfg.index1_len > 0 and self.cfg.index2_len > 0
        hm = self.L.orc_hopmap_new(Lb) if hop_on else None
        try:
            rc = self.L.orc_demux_batch(self._m, reads.ctypes.data, n,
                                        idx.ctypes.data if idx is not None else None,
                                        dist.ctypes.data if dist is not None else None,
                                        per_sample.ctypes.data, C.byref(total), hm, threads, int(use_cache),
                                        C.byref(secs))
            assert rc == 0
            hops: Dict[bytes, int] = {}
            if hop_on:
                k = self.L.orc_hopmap_len(hm)
                keys = np.zeros((max(k, 1), Lb), dtype=np.uint8)
                cnt = np.zeros(max(k, 1), dtype=np.uint64)
                self.L.orc_hopmap_export(hm, keys.ctypes.data, cnt.ctypes.data)
                hops = {keys[i].tobytes(): int(cnt[i]) for i in range(k)}
        finally:
            if hm:
                self.L.orc_hopmap_free(hm)
        return {"idx": idx, "dist": dist, "per_sample": per_sample, "total_templates": int(total.value),
                "hops": hops, "seconds": secs.value}


def unmatched_counts(reads: np.ndarray, idx: np.ndarray, undetermined: int) -> Dict[bytes, int]:
    """Counts of concatenated barcodes routed to Undetermined (input of metrics.rs:151-166 before
    delimiting/down-sizing). Host-side helper used to check most_frequent_unmatched parity."""
    sel = reads[idx == undetermined]
    if sel.shape[0] == 0:
        return {}
    u, c = np.unique(sel, axis=0, return_counts=True)
    return {u[i].tobytes(): int(c[i]) for i in range(u.shape[0])}


# ------------------------------------------------------------------------------------------------
# The stages either side of the assignment, restated with numpy / plain Python.
# TEST INFRASTRUCTURE ONLY, like the rest of this file.
# ------------------------------------------------------------------------------------------------
class UnmatchedCounterLiteral:
    """metrics.rs:131-190, literally: insert() down-sizes when the map already holds max_counter_size keys
    (metrics.rs:151-160), downsize() keeps the downsize_to most frequent (166-175; ties in hash order in the
    reference, in byte order here), to_file() writes the n most frequent in descending count (177-190)."""

    def __init__(self, max_counter_size: int = 5_000_000, downsize_to: int = 5_000):
        self.unmatched_counter: Dict[bytes, int] = {}
        self.max_counter_size, self.downsize_to = max_counter_size, downsize_to

    def insert(self, barcode: bytes) -> None:
        if len(self.unmatched_counter) == self.max_counter_size:
            self.downsize()
        self.unmatched_counter[barcode] = self.unmatched_counter.get(barcode, 0) + 1

    def downsize(self) -> None:
        top = sorted(self.unmatched_counter.items(), key=lambda kv: (-kv[1], kv[0]))[:self.downsize_to]
        self.unmatched_counter = dict(top)

    def top(self, n: int):
        return sorted(self.unmatched_counter.items(), key=lambda kv: (-kv[1], kv[0]))[:n]


def extract_sample_barcodes(sources: Sequence[np.ndarray], structures) -> np.ndarray:
    """demux.rs:360-417 + 504-522: the B segments of every FASTQ's read, FASTQ order then segment order,
    concatenated.  sources[f] = [n, stride_f] bases; structures[f] = list of (offset, length|None, kind)."""
    parts = []
    for bases, segs in zip(sources, structures):
        for (off, length, kind) in segs:
            if kind == "B":
                parts.append(bases[:, off:(bases.shape[1] if length is None else off + length)])
    return np.ascontiguousarray(np.concatenate(parts, axis=1))


def base_qual_counts(quals: np.ndarray, lens: Optional[np.ndarray], segments, idx: np.ndarray,
                     num_samples: int) -> Tuple[np.ndarray, int]:
    """BaseQualCounter::update (metrics.rs:499-518) for every segment of one FASTQ's reads, summed per assigned
    sample (demux.rs:582-583).  q = byte - 33 in u8 arithmetic (wraps below 33, as `q - 33` on a u8 does in a
    release build).  Returns ([(S+1), 5] = q20, q30, index_bases_qual_sum, index_bases_total_seen, bases; number
    of reads too short for their fixed segments, which add nothing -- the reference fails on those)."""
    n, stride = quals.shape
    out = np.zeros((num_samples + 1, 5), dtype=np.uint64)
    lens = np.full(n, stride, dtype=np.int64) if lens is None else np.minimum(lens.astype(np.int64), stride)
    need = max([off + (length or 0) for (off, length, kind) in segments] + [0])
    ok = lens >= need
    q = (quals.astype(np.uint8) - np.uint8(33)).astype(np.uint8)  # wrapping
    pos = np.arange(stride)[None, :]
    bins = idx.astype(np.int64)
    for (off, length, kind) in segments:
        end = lens if length is None else np.full(n, off + length, dtype=np.int64)
        inside = (pos >= off) & (pos < end[:, None]) & ok[:, None]
        cnt = inside.sum(axis=1)
        if kind == "B":
            np.add.at(out[:, 3], bins, cnt.astype(np.uint64))
            np.add.at(out[:, 2], bins, (q * inside).sum(axis=1, dtype=np.uint64))
        elif kind == "T":
            np.add.at(out[:, 4], bins, cnt.astype(np.uint64))
            np.add.at(out[:, 0], bins, ((q >= 20) & inside).sum(axis=1).astype(np.uint64))
            np.add.at(out[:, 1], bins, ((q >= 30) & inside).sum(axis=1).astype(np.uint64))
    return out, int((~ok).sum())


def route_reads(idx: np.ndarray, num_samples: int) -> Tuple[np.ndarray, np.ndarray]:
    """What DemuxedGroup::per_sample_reads holds (demux.rs:578,637-645): the reads of each sample in input
    order.  Returns (order, offsets): order[offsets[s]:offsets[s+1]] = read ids of sample s; S+2 offsets."""
    order = np.argsort(idx, kind="stable").astype(np.uint32)
    counts = np.bincount(idx.astype(np.int64), minlength=num_samples + 1)
    offsets = np.zeros(num_samples + 2, dtype=np.uint64)
    offsets[1:] = np.cumsum(counts)
    return order, offsets


class LiteralDemultiplexer:
    """Demultiplexer::demultiplex (demux.rs:445-584) for templates that passed the header filters, read by read in
    plain Python (small inputs): process_segments over every FASTQ (360-417: extraction, BaseQualCounter), the
    concatenated sample barcode (504-522), the N gate and the matcher (524-538), update_with_match, hop check and
    unmatched collection (541-556), routing (578), total_templates and the quality counters of the template's
    sample (581-583).  fastqs = [(bases [n, stride], quals [n, stride], lens [n] or None)], structures[f] = list
    of (offset, length | None, kind)."""

    def __init__(self, cfg: Config, samples: Sequence[bytes]):
        self.demux = LiteralDemux(cfg, samples)
        self.cfg = cfg

    def run(self, fastqs, structures):
        S = self.cfg.num_samples
        n = fastqs[0][0].shape[0]
        qual = np.zeros((S + 1, 5), dtype=np.uint64)
        per_sample_reads = [[] for _ in range(S + 1)]
        barcodes, counters = [], []
        for r in range(n):
            bc = b""
            counter = [0, 0, 0, 0, 0]  # q20, q30, index sum, index seen, bases
            for (bases, quals, lens), segs in zip(fastqs, structures):
                ln = bases.shape[1] if lens is None else int(lens[r])
                for (off, length, kind) in segs:
                    end = ln if length is None else off + length
                    sb, sq = bases[r, off:end], quals[r, off:end]
                    if kind == "B":
                        bc += sb.tobytes()
                        counter[3] += len(sq)
                        counter[2] += int(sum((int(q) - 33) & 0xFF for q in sq))
                    elif kind == "T":
                        counter[4] += len(sq)
                        for q in sq:
                            q = (int(q) - 33) & 0xFF
                            if q >= 30:
                                counter[1] += 1
                                counter[0] += 1
                            elif q >= 20:
                                counter[0] += 1
            barcodes.append(bc)
            counters.append(counter)
        res = self.demux.run(barcodes)
        for r in range(n):  # routing (578) and the template's quality counters added to its sample (582-583)
            b = int(res["idx"][r])
            per_sample_reads[b].append(r)
            qual[b] += np.array(counters[r], dtype=np.uint64)
        res["per_sample_reads"] = per_sample_reads
        res["base_qual"] = qual
        res["barcodes"] = barcodes
        return res


def demultiplex_batch(cfg: Config, samples: Sequence[bytes], fastqs, structures):
    """The same stage composed from the vectorised restatements above (any size): returns idx, dist, per_sample,
    hops, unmatched, base_qual [(S+1), 5], order, offsets."""
    S = cfg.num_samples
    barcodes = extract_sample_barcodes([f[0] for f in fastqs], structures)
    res = COracle(cfg, [bytes(s) for s in samples]).run(barcodes, threads=4)
    qual = np.zeros((S + 1, 5), dtype=np.uint64)
    for (bases, quals, lens), segs in zip(fastqs, structures):
        qual += base_qual_counts(quals, lens, segs, res["idx"], S)[0]
    res["base_qual"] = qual
    res["unmatched"] = unmatched_counts(barcodes, res["idx"], S)
    res["order"], res["offsets"] = route_reads(res["idx"], S)
    res["barcodes"] = barcodes
    return res


def compact_records(reads: np.ndarray) -> np.ndarray:
    """The compact read format of include/sgdx.h ("Compact records"), stated with numpy: one 64-bit word per
    read of at most 21 bases = lo plane | hi plane << 21 | invalid plane << 42, code = (ascii >> 1) & 3
    (A0 C1 T2 G3); at a position that is not A/C/G/T the invalid bit is set, hi = 0, and lo = 0 for 'N'
    (the byte demux.rs:78-80 counts as a no-call) or 1 for any other byte (which mismatches every sample,
    matcher.rs:335-348).  Checker of sgdx_pack_compact_host / sgdx_pack_compact_device."""
    reads = np.asarray(reads, dtype=np.uint8)
    n, L = reads.shape
    assert L <= 21
    a = reads.astype(np.uint64)
    ok = np.isin(reads, np.frombuffer(b"ACGT", dtype=np.uint8))
    j = np.arange(L, dtype=np.uint64)
    lo = np.where(ok, (a >> np.uint64(1)) & np.uint64(1), (reads != ord("N")).astype(np.uint64))
    hi = np.where(ok, (a >> np.uint64(2)) & np.uint64(1), np.uint64(0))
    inv = (~ok).astype(np.uint64)
    if n == 0:
        return np.empty(0, dtype=np.uint64)
    return ((lo << j).sum(axis=1) | ((hi << j).sum(axis=1) << np.uint64(21)) |
            ((inv << j).sum(axis=1) << np.uint64(42))).astype(np.uint64)


def dense_records(reads: np.ndarray):
    """The dense records of include/sgdx.h, stated with numpy: per read ceil(2L/8) bytes, bits 0..L-1 the lo plane, bits
    L..2L-1 the hi plane (code = (ascii >> 1) & 3), little endian; reads with a byte outside A/C/G/T are exceptions
    (index, compact record) and their dense bytes are unspecified.  Returns (dense [n, B] uint8, exception indices,
    exception compact records, mask of the plain reads).  Checker of sgdx_pack_dense_host."""
    reads = np.asarray(reads, dtype=np.uint8)
    n, L = reads.shape
    assert L <= 32  # (above 21 bases the exceptions are sgdx_read records, see packed_records)
    B = (2 * L + 7) // 8
    ok = np.isin(reads, np.frombuffer(b"ACGT", dtype=np.uint8))
    plain = ok.all(axis=1)
    a = reads.astype(np.uint64)
    j = np.arange(L, dtype=np.uint64)
    lo = (np.where(ok, (a >> np.uint64(1)) & np.uint64(1), np.uint64(0)) << j).sum(axis=1).astype(np.uint64) if n else np.zeros(0, np.uint64)
    hi = (np.where(ok, (a >> np.uint64(2)) & np.uint64(1), np.uint64(0)) << j).sum(axis=1).astype(np.uint64) if n else np.zeros(0, np.uint64)
    x = lo | (hi << np.uint64(L))
    dense = np.zeros((n, B), dtype=np.uint8)
    for b in range(B):
        dense[:, b] = ((x >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8)
    exc = np.nonzero(~plain)[0].astype(np.uint32)
    if L <= 21:
        recs = compact_records(reads[exc]) if len(exc) else np.zeros(0, np.uint64)
    else:
        recs = packed_records(reads[exc]) if len(exc) else np.zeros((0, 4), np.uint32)
    return dense, exc, recs, plain


def packed_records(reads: np.ndarray) -> np.ndarray:
    """The 16-byte sgdx_read records of include/sgdx.h, stated with numpy: (n, 4) uint32 {lo, hi, nmask, xmask},
    bit j = base j; code = (ascii >> 1) & 3 at A/C/G/T positions (planes zero elsewhere), nmask = byte == 'N',
    xmask = any other byte.  Checker of sgdx_pack_host / sgdx_pack_device."""
    reads = np.asarray(reads, dtype=np.uint8)
    n, L = reads.shape
    assert L <= 32
    a = reads.astype(np.uint64)
    ok = np.isin(reads, np.frombuffer(b"ACGT", dtype=np.uint8))
    isn = reads == ord("N")
    j = np.arange(L, dtype=np.uint64)
    out = np.zeros((n, 4), dtype=np.uint32)
    if n:
        out[:, 0] = (np.where(ok, (a >> np.uint64(1)) & np.uint64(1), np.uint64(0)) << j).sum(axis=1)
        out[:, 1] = (np.where(ok, (a >> np.uint64(2)) & np.uint64(1), np.uint64(0)) << j).sum(axis=1)
        out[:, 2] = (isn.astype(np.uint64) << j).sum(axis=1)
        out[:, 3] = ((~ok & ~isn).astype(np.uint64) << j).sum(axis=1)
    return out


# ------------------------------------------------------------------------------------------------
# seeded synthetic workloads without the product library: oracle/liboracle_synth.so is built
# from the generator's shared source (singular-demux_b200/csrc/sgdx_synth_core.h) by oracle/Makefile
# ------------------------------------------------------------------------------------------------
_SYNTH = None


def synth_lib() -> C.CDLL:
    global _SYNTH
    if _SYNTH is None:
        out = os.path.join(_HERE, "liboracle_synth.so")
        src = os.path.join(_HERE, "synth_host.cpp")
        core = os.path.join(os.path.dirname(_HERE), "singular-demux_b200", "csrc", "sgdx_synth_core.h")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(core)):
            subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-o", out, src, "-lpthread"], check=True, cwd=_HERE)
        L = C.CDLL(out)
        L.orc_synth_table.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_synth_reads_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        _SYNTH = L
    return _SYNTH


def synth_table(w) -> np.ndarray:
    """Expected-barcode table of a Workload (singular_demux_b200.synth), drawn by the oracle-side generator."""
    out = np.zeros((w.num_samples, w.barcode_len), dtype=np.uint8)
    if synth_lib().orc_synth_table(w.table_seed, w.num_samples, w.barcode_len, w.index1_len, w.min_dist, out.ctypes.data):
        raise RuntimeError("could not draw the table")
    return out


def synth_reads(w, first: int, n: int, table: np.ndarray, threads: int = 8) -> np.ndarray:
    out = np.empty((n, w.barcode_len), dtype=np.uint8)
    p = w.params()
    tb = np.ascontiguousarray(table, dtype=np.uint8)
    if synth_lib().orc_synth_reads_host(C.byref(p), tb.ctypes.data, first, n, out.ctypes.data, threads):
        raise RuntimeError("synthetic read generation failed")
    return out

