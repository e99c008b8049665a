"""Breakend homology of an indel: how far it can slide without changing the edit distance.

Host-side mirror of lib/rust-vc-utils/src/indel_breakend_homology.rs:24-73 (get_indel_breakend_homology_info),
which the reference applies to every indel / breakpoint pulled out of a contig CIGAR
(src/refine_sv/align/mod.rs:85-102, src/refine_sv/align/segment_alignment.rs:318-336).
Ranges are half-open (start, end) pairs, zero-based, starting at the first position the indel touches:
2M1D2M -> ref (2, 3), read (2, 2); 2M1I2M -> ref (2, 2), read (2, 3).
"""
from __future__ import annotations

from typing import Tuple

Range = Tuple[int, int]


def get_indel_breakend_homology_info(ref_seq: bytes, ref_range: Range, read_seq: bytes,
                                     read_range: Range) -> Tuple[Range, bytes]:
    """Returns ((-left, right) offsets the indel can move by, homology sequence left part + right part).

    Sliding left compares the bases just before the two range ENDS, sliding right the bases at the two range
    STARTS; both stop at the first mismatch or at a sequence edge. Flanks are assumed to match perfectly.
    """
    r0, r1 = ref_range
    q0, q1 = read_range
    left = 0
    limit = min(r0, q0)
    while left < limit and ref_seq[r1 - left - 1] == read_seq[q1 - left - 1]:
        left += 1
    right = 0
    limit = min(len(ref_seq) - r1, len(read_seq) - q1)
    while right < limit and ref_seq[r0 + right] == read_seq[q0 + right]:
        right += 1
    hom = bytes(ref_seq[r1 - left:r1]) + bytes(ref_seq[r0:r0 + right])
    return (-left, right), hom
