"""Host-side mirror of the reference's alignment result type.

Follows src/simple_alignment.rs:6-72 (SimpleAlignment, get_reverse, clip_alignment_ref_edges) and the cigar
helpers it uses from lib/rust-vc-utils/src/bam_utils/cigar/{mod.rs:41-47,204-232, clip_alignment.rs:15-90}.
Cigar elements are (op, length) tuples with op one of "MIDNSHP=X" (htslib order); the C ABI carries them as
htslib u32 (length << 4 | op index).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

CIGAR_OPS = "MIDNSHP=X"
Cigar = Tuple[str, int]


def cigar_from_u32(words) -> List[Cigar]:
    return [(CIGAR_OPS[int(w) & 0xF], int(w) >> 4) for w in words]


def cigar_to_u32(cigar: List[Cigar]) -> List[int]:
    return [(n << 4) | CIGAR_OPS.index(op) for op, n in cigar]


def cigar_to_string(cigar: List[Cigar]) -> str:
    return "".join(f"{n}{op}" for op, n in cigar)


def cigar_from_string(s: str) -> List[Cigar]:
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((ch, int(num)))
            num = ""
    return out


def get_cigarseg_ref_offset(c: Cigar) -> int:
    return c[1] if c[0] in "DN X=M".replace(" ", "") else 0


def get_cigarseg_read_offset(c: Cigar, ignore_hard_clip: bool = True) -> int:
    op, n = c
    if op in "MI=XS":
        return n
    if op == "H" and not ignore_hard_clip:
        return n
    return 0


def get_cigar_ref_offset(cigar: List[Cigar]) -> int:
    return sum(get_cigarseg_ref_offset(c) for c in cigar)


def get_cigar_read_offset(cigar: List[Cigar], ignore_hard_clip: bool = True) -> int:
    return sum(get_cigarseg_read_offset(c, ignore_hard_clip) for c in cigar)


def compress_cigar(cigar: List[Cigar]) -> List[Cigar]:
    out: List[Cigar] = []
    for op, n in cigar:
        if n == 0:
            continue
        if out and out[-1][0] == op:
            out[-1] = (op, out[-1][1] + n)
        else:
            out.append((op, n))
    return out


def clip_alignment_ref_start(cigar_in: List[Cigar], min_left_ref_clip: int):
    ref_pos = 0
    out: List[Cigar] = []
    shift = 0
    for c in cigar_in:
        op, n = c
        if op in "DN":
            if ref_pos <= min_left_ref_clip:
                shift += n
            else:
                out.append(c)
        elif op == "I":
            out.append(("S", n) if ref_pos < min_left_ref_clip else c)
        elif op in "MX=":
            if ref_pos < min_left_ref_clip:
                remaining = min_left_ref_clip - shift
                match_size = max(n - remaining, 0)
                clip_size = n - match_size
                out.append(("S", clip_size))
                if match_size > 0:
                    out.append((op, match_size))
                shift += clip_size
            else:
                out.append(c)
        else:
            out.append(c)
        ref_pos += get_cigarseg_ref_offset(c)
    return out, shift


def cigar_clip_alignment_ref_edges(cigar_in: List[Cigar], min_left_ref_clip: int, min_right_ref_clip: int):
    right, _ = clip_alignment_ref_start(list(reversed(cigar_in)), min_right_ref_clip)
    right.reverse()
    clipped, ref_shift = clip_alignment_ref_start(right, min_left_ref_clip)
    return compress_cigar(clipped), ref_shift


@dataclass
class SimpleAlignment:
    ref_offset: int = 0
    cigar: List[Cigar] = field(default_factory=list)

    def get_reverse(self, ref_size: int) -> "SimpleAlignment":
        ref_end_offset = self.ref_offset + get_cigar_ref_offset(self.cigar)
        assert ref_end_offset <= ref_size
        return SimpleAlignment(ref_size - ref_end_offset, list(reversed(self.cigar)))

    def __repr__(self) -> str:
        return (f"offset: {self.ref_offset} cigar_ref_len: {get_cigar_ref_offset(self.cigar)} "
                f"cigar_read_len: {get_cigar_read_offset(self.cigar, True)} cigar: {cigar_to_string(self.cigar)}")


def clip_alignment_ref_edges(alignment: SimpleAlignment, ref_size: int, min_left_ref_clip: int,
                             min_right_ref_clip: int) -> SimpleAlignment:
    ref_end_offset = alignment.ref_offset + get_cigar_ref_offset(alignment.cigar)
    right_ref_offset = ref_size - ref_end_offset
    clipped, shift = cigar_clip_alignment_ref_edges(alignment.cigar, min_left_ref_clip - alignment.ref_offset,
                                                    min_right_ref_clip - right_ref_offset)
    return SimpleAlignment(alignment.ref_offset + shift, clipped)
