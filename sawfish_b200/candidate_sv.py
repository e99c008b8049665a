"""Pieces of the `candidate.sv.bcf` side of the discover -> joint-call hand-off that do not need a BCF parser.

Reference: src/joint_call/get_refined_svs.rs.  Built so far: `parse_bnd_alt_allele` (:206-287), pinned by the reference's
two unit tests (:926-976).  The BCF2 record reader itself (typed INFO values, string / contig dictionaries, :334-606) is
not built; `sawfish_b200/contig_bam.py` covers the contig half of the hand-off.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

LEFT_ANCHOR, RIGHT_ANCHOR = "LeftAnchor", "RightAnchor"  # crate::breakpoint::BreakendDirection


@dataclass
class BndAltInfo:  # :196-204
    breakend1_direction: str
    breakend2_chrom_index: int
    breakend2_pos: int
    breakend2_direction: str
    insert_seq: bytes


def parse_bnd_alt_allele(label_to_index: Dict[str, int], alt_allele: bytes, homlen: int) -> BndAltInfo:
    """Parse a VCF breakend alt allele such as b"TCT]chr1:500]" (:206-287).  `homlen` corrects the remote breakend to
    its left-shifted position when both breakends have the same orientation (inverted breakpoint)."""
    s = alt_allele.decode()
    if "]" in s:
        breakend2_direction, split_char = LEFT_ANCHOR, "]"
    elif "[" in s:
        breakend2_direction, split_char = RIGHT_ANCHOR, "["
    else:
        raise ValueError(f"Invalid BND alt alt allele '{s}'")
    words = s.split(split_char)
    assert len(words) == 3, f"Unexpected BND alt allele format `{s}`"
    mate = words[1].rsplit(":", 1)  # chromosome names may contain ':' (HLA contigs): split at the last one
    assert len(mate) == 2, f"Unexpected BND alt allele format `{s}`"
    breakend2_chrom_index = label_to_index[mate[0]]
    breakend2_raw_pos = int(mate[1]) - 1
    if words[0] and not words[2]:
        breakend1_direction = LEFT_ANCHOR
    elif not words[0] and words[2]:
        breakend1_direction = RIGHT_ANCHOR
    else:
        raise ValueError(f"Invalid BND alt alt allele '{s}'")
    insert_seq = words[0].encode()[1:] if breakend1_direction == LEFT_ANCHOR else words[2].encode()[:-1]
    pos = breakend2_raw_pos
    if breakend1_direction == breakend2_direction:
        pos -= homlen
    if breakend2_direction == RIGHT_ANCHOR:
        pos -= 1
    return BndAltInfo(breakend1_direction, breakend2_chrom_index, pos, breakend2_direction, insert_seq)
