"""Read structures as the reference takes them on the command line (`--read-structures 8B+T 10B`,
opts.rs:84-92; crate read-structure 0.1.0, used at demux.rs:380-387): a sequence of <length><kind> segments,
kinds T (template) B (sample barcode) M (molecular barcode) S (skip); only the last segment may have the
indefinite length `+`.  Only what the device stages need: offsets, lengths, kinds.
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass
from typing import List, Optional, Sequence

from ._lib import SEG_TO_END, Segment

KINDS = "TBMS"


@dataclass(frozen=True)
class ReadSegment:
    offset: int
    length: Optional[int]   # None = `+`, to the end of the read
    kind: str

    def end(self, read_len: int) -> int:
        return read_len if self.length is None else self.offset + self.length


@dataclass(frozen=True)
class ReadStructure:
    segments: tuple

    @staticmethod
    def from_str(s: str) -> "ReadStructure":
        s = s.strip().upper()
        toks = re.findall(r"(\d+|\+)([A-Z])", s)
        if not toks or "".join(a + b for a, b in toks) != s:
            raise ValueError(f"Read structure {s!r} is not a sequence of <length><kind> segments")
        segs: List[ReadSegment] = []
        off = 0
        for i, (n, k) in enumerate(toks):
            if k not in KINDS:
                raise ValueError(f"Read structure {s!r}: unknown segment kind {k!r}")
            if n == "+":
                if i != len(toks) - 1:
                    raise ValueError(f"Read structure {s!r}: only the last segment may have indefinite length")
                segs.append(ReadSegment(off, None, k))
            else:
                if int(n) == 0:
                    raise ValueError(f"Read structure {s!r}: zero-length segment")
                segs.append(ReadSegment(off, int(n), k))
                off += int(n)
        return ReadStructure(tuple(segs))

    def sample_barcode_length(self, read_len: Optional[int] = None) -> int:
        """Sum of the B-segment lengths (metrics.rs:588-595)."""
        tot = 0
        for g in self.segments:
            if g.kind == "B":
                if g.length is None:
                    if read_len is None:
                        raise ValueError("indefinite-length sample barcode segment needs the read length")
                    tot += max(read_len - g.offset, 0)
                else:
                    tot += g.length
        return tot

    def min_length(self) -> int:
        return max([g.offset + (g.length or 0) for g in self.segments] + [0])


def barcode_segment_lengths(structures: Sequence[ReadStructure], read_lens: Optional[Sequence[int]] = None) -> List[int]:
    """Lengths of the sample-barcode segments in FASTQ order, then segment order (demux.rs:504-522): the pieces the
    delimited barcode joins with `+` (demux.rs:56-62).  read_lens[f] is needed for a trailing `+B`."""
    out: List[int] = []
    for f, rs in enumerate(structures):
        for g in rs.segments:
            if g.kind == "B":
                if g.length is None:
                    if read_lens is None:
                        raise ValueError("indefinite-length sample barcode segment needs the read length")
                    out.append(max(read_lens[f] - g.offset, 0))
                else:
                    out.append(g.length)
    return out


def c_segments(structures: Sequence[ReadStructure]):
    """sgdx_segment array of several FASTQs' structures (source = position in `structures`)."""
    rows = [(src, g) for src, rs in enumerate(structures) for g in rs.segments]
    arr = (Segment * len(rows))()
    for i, (src, g) in enumerate(rows):
        arr[i] = Segment(src, g.offset, SEG_TO_END if g.length is None else g.length, ord(g.kind))
    return arr, len(rows)
