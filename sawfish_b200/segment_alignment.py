"""Split a contig-vs-alt-haplotype alignment (segment 1 + N spacer + segment 2) into its two segment alignments.

Host-side mirror of src/refine_sv/align/segment_alignment.rs:26-357: the consumer of K2's CIGAR for two-region
(breakpoint) candidates, called from src/refine_sv/align/mod.rs:1340-1420 right after the large-indel alignment.
What it observes of the CIGAR: an `=` run of >= min_assembly_edge_anchor on each side of the spacer, no op ending
inside the spacer, the spacer crossed by an op that consumes no contig bases, and (unless the left segment is pinned)
no op ending exactly on the first spacer position.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

from .indel_breakend_homology import get_indel_breakend_homology_info
from .simple_alignment import (Cigar, SimpleAlignment, get_cigar_ref_offset, get_cigarseg_read_offset,
                               get_cigarseg_ref_offset)


@dataclass
class TwoRegionAltHapInfo:
    """The fields of the reference's TwoRegionAltHapInfo (src/refine_sv/align/mod.rs:640-700) this step reads."""
    alt_hap_seq: bytes
    spacer_range: Tuple[int, int]
    left_boundary_pin: bool = False
    right_boundary_pin: bool = False
    extended_alt_hap_seq: Optional[bytes] = None

    def __post_init__(self):
        if self.extended_alt_hap_seq is None:
            self.extended_alt_hap_seq = self.alt_hap_seq


@dataclass
class AltHapLeftRightComponentAlignmentInfo:
    left_alignment: SimpleAlignment
    right_alignment: SimpleAlignment
    left_right_breakend_homology_range: Tuple[int, int]
    left_right_breakend_homology_seq: bytes


def does_alignment_support_alt_hap(min_assembly_edge_anchor: int, alignment: SimpleAlignment, boundary1: int,
                                   boundary2: int) -> bool:
    """An `=` run with min_assembly_edge_anchor bases strictly before the spacer, and one with as many at/after its end
    (segment_alignment.rs:26-57)."""
    left = right = False
    pos = alignment.ref_offset
    for c in alignment.cigar:
        op, n = c
        if op == "=" and n >= min_assembly_edge_anchor:
            left = left or pos + min_assembly_edge_anchor < boundary1
            right = right or pos + n - min_assembly_edge_anchor >= boundary2
        pos += get_cigarseg_ref_offset(c)
    return left and right


def _left_segment(alignment: SimpleAlignment, boundary1: int, pinned: bool):
    """CIGAR of the part before the spacer; everything from the op that crosses into the spacer on is soft-clipped."""
    ops = alignment.cigar
    ends, pos = [], alignment.ref_offset
    for c in ops:
        pos += get_cigarseg_ref_offset(c)
        ends.append(pos)
    cross = next((i for i, e in enumerate(ends) if e > boundary1), len(ops))
    if not pinned and any(e == boundary1 for e in ends[:cross]):
        return None  # an op may end on the first spacer position only when the left segment is pinned
    kept: List[Cigar] = list(ops[:max(cross - 1, 0)]) if cross < len(ops) else list(ops[:-1])
    clip = 0
    if 0 < cross < len(ops):
        before = ops[cross - 1]
        if get_cigarseg_ref_offset(before) > 0:
            kept.append(before)
        else:
            clip += get_cigarseg_read_offset(before, True)  # an insertion abutting the breakpoint goes to the clip
    clip += sum(get_cigarseg_read_offset(c, True) for c in ops[cross + 1:])  # (the crossing op itself is not counted)
    if clip > 0:
        kept.append(("S", clip))
    return kept, clip


def _right_segment(alignment: SimpleAlignment, boundary1: int, boundary2: int):
    """CIGAR of the part after the spacer, its offset in the right segment, and the contig bases clipped before it."""
    pos = alignment.ref_offset
    inside = False
    offset = None
    cigar: List[Cigar] = []
    clip = 0
    for c in alignment.cigar:
        ref_n, read_n = get_cigarseg_ref_offset(c), get_cigarseg_read_offset(c, True)
        if not inside:
            if pos + ref_n >= boundary2:
                if pos < boundary1 and read_n > 0:
                    return None  # a match block spans the spacer
                inside = True
            else:
                if pos + ref_n > boundary1:
                    return None  # an op ends inside the spacer
                clip += read_n
        else:
            if offset is None and read_n > 0:
                assert pos >= boundary2
                offset = pos - boundary2
            if not cigar:
                if ref_n == 0:
                    clip += read_n
                else:
                    if clip > 0:
                        cigar.append(("S", clip))
                    cigar.append(c)
            else:
                cigar.append(c)
        pos += ref_n
    assert cigar and offset is not None
    return cigar, offset, clip


def transform_alt_hap_alignment_into_left_right_components(
        min_assembly_edge_anchor: int, alignment: SimpleAlignment, alt_hap_info: TwoRegionAltHapInfo,
        contig: bytes) -> Optional[AltHapLeftRightComponentAlignmentInfo]:
    """segment_alignment.rs:125-357. None when the alignment does not support the alt haplotype."""
    b1, b2 = alt_hap_info.spacer_range
    if not does_alignment_support_alt_hap(min_assembly_edge_anchor, alignment, b1, b2):
        return None
    left = _left_segment(alignment, b1, alt_hap_info.left_boundary_pin)
    if left is None:
        return None
    left_cigar, left_clip = left
    right = _right_segment(alignment, b1, b2)
    if right is None:
        return None
    right_cigar, right_offset, right_clip = right
    # the breakpoint shows as an indel on the alt haplotype: its span there and on the contig
    sv_alt_hap_range = (alignment.ref_offset + get_cigar_ref_offset(left_cigar), b2 + right_offset)
    sv_contig_range = (len(contig) - left_clip, right_clip)
    ext = alt_hap_info.extended_alt_hap_seq
    assert len(ext) >= len(alt_hap_info.alt_hap_seq)
    grow = len(ext) - len(alt_hap_info.alt_hap_seq)
    hom_range, hom_seq = get_indel_breakend_homology_info(ext, (sv_alt_hap_range[0], sv_alt_hap_range[1] + grow), contig,
                                                          sv_contig_range)
    return AltHapLeftRightComponentAlignmentInfo(SimpleAlignment(alignment.ref_offset, left_cigar),
                                                 SimpleAlignment(right_offset, right_cigar), hom_range, hom_seq)


def collapse_indels_at_breakpoint(alt_hap_info: TwoRegionAltHapInfo, alignment: SimpleAlignment,
                                  min_assembly_collapse_anchor: int = 4) -> None:
    """Fold every edit between the last >= 4-base `=` anchor before the spacer and the first one after it into a single
    insertion + deletion at the breakpoint, in place (src/refine_sv/align/mod.rs:1416-1518). Applied to K2's CIGAR
    before the left/right split so that breakpoint insertions are represented the same way whatever the aligner's
    tie-breaks did with small edits next to the spacer."""
    b1, b2 = alt_hap_info.spacer_range
    anchored1 = anchored2 = False
    first_edit = last_edit = None
    pos = alignment.ref_offset
    for i, (op, n) in enumerate(alignment.cigar):
        if op == "=":
            if n >= min_assembly_collapse_anchor:
                if pos + min_assembly_collapse_anchor < b1:
                    anchored1, first_edit = True, None  # a later anchor in segment 1 supersedes earlier edits
                if pos + n - min_assembly_collapse_anchor >= b2:
                    anchored2 = True
                    break
        elif op in "DXI":
            if anchored1 and first_edit is None:
                first_edit = i
            last_edit = i
        pos += get_cigarseg_ref_offset((op, n))
    if not (anchored1 and anchored2) or first_edit is None or last_edit is None:
        return
    assert last_edit >= first_edit
    if first_edit == last_edit and alignment.cigar[first_edit][0] == "D":
        return  # already a clean deletion across the spacer
    inner = alignment.cigar[first_edit:last_edit + 1]
    merged: List[Cigar] = []
    ins = sum(get_cigarseg_read_offset(c, True) for c in inner)
    dele = sum(get_cigarseg_ref_offset(c) for c in inner)
    if ins > 0:
        merged.append(("I", ins))
    if dele > 0:
        merged.append(("D", dele))
    tail = alignment.cigar[last_edit + 1:]
    # the reference emits the merged pair when it meets the first op after the block, so a block that ends the CIGAR is dropped
    alignment.cigar = alignment.cigar[:first_edit] + (merged + tail if tail else [])
