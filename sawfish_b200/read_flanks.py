"""Read side of score_sv's pair enumeration: which reads qualify, and which stretch of each read becomes the `text` of an
alignment pair.  Host-side mirror (no GPU work here; the pairs these ranges cut go to K1 through `score_sv.PairBatchBuilder`).

Reference:
  get_gap_compressed_identity_from_alignment / get_final_gci   lib/rust-vc-utils/src/bam_utils/cigar/score_alignment.rs:68-128
  get_gap_compressed_identity_from_cigar_segment_range          src/bam_utils.rs:21-55
  get_alignment_closest_to_target_ref_pos                       src/bam_utils.rs:95-150
  translate_ref_range_to_hardclipped_read_range                 src/bam_utils.rs:180-243
  get_alignment_ref_segments                                    src/bam_utils.rs:339-371
  get_read_breakend_flank_range                                 src/score_sv/mod.rs:321-432
  the read filter of get_read_support_scores_from_segment       src/score_sv/mod.rs:784-797 (mapq, gap-compressed identity)
All of the reference's unit tests for these functions (src/bam_utils.rs:378-640) are replayed in tests/test_read_flanks.py.
Cigar elements are (op, length) tuples as in simple_alignment.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from .simple_alignment import Cigar

LEFT_ANCHOR, RIGHT_ANCHOR = "LeftAnchor", "RightAnchor"
ANY, TARGET_OR_LOWER, TARGET_OR_HIGHER, TARGET = "Any", "TargetOrLower", "TargetOrHigher", "Target"  # TargetMatchType

READ_SUPPORT_FLANK_SIZE = 500                    # ScoreSVSettings::new, src/score_sv/mod.rs:2047-2055
MIN_READ_SUPPORT_FLANK_SIZE = 100
READ_SUPPORT_ANCHOR_MIN_EDGE_DISTANCE = 200
REF_TO_READ_ANCHOR_OFFSET = 10                   # :328


def _ref_advance(c: Cigar) -> int:   # update_ref_pos
    return c[1] if c[0] in "MDN=X" else 0


def _read_advance(c: Cigar, ignore_hard_clip: bool = True) -> int:   # update_read_pos
    if c[0] in "MIS=X":
        return c[1]
    if c[0] == "H" and not ignore_hard_clip:
        return c[1]
    return 0


def get_gap_compressed_identity_from_alignment(ref_pos: int, cigar: Sequence[Cigar], read_seq: bytes, ref_seq: bytes,
                                               ignore_hard_clip: bool = True) -> float:
    """Matches / (matches + mismatches + gap EVENTS): every insertion or deletion counts once, whatever its length."""
    mismatch_events = match_bases = read_pos = 0
    for c in cigar:
        op, n = c
        if op in "ID":
            mismatch_events += 1
        elif op == "X":
            mismatch_events += n
        elif op == "=":
            match_bases += n
        elif op == "M":
            for o in range(n):
                if ref_seq[ref_pos + o] == read_seq[read_pos + o]:
                    match_bases += 1
                else:
                    mismatch_events += 1
        ref_pos += _ref_advance(c)
        read_pos += _read_advance(c, ignore_hard_clip)
    return 1.0 if match_bases + mismatch_events == 0 else match_bases / (match_bases + mismatch_events)


def get_gap_compressed_identity_from_cigar_segment_range(ref_pos: int, read_seq: bytes, cigar: Sequence[Cigar], chrom_seq: bytes,
                                                         cigar_start_index: int, cigar_end_index: int) -> float:
    assert cigar_start_index < cigar_end_index <= len(cigar)
    read_offset = 0
    for c in cigar[:cigar_start_index]:
        ref_pos += _ref_advance(c)
        read_offset += _read_advance(c, True)
    return get_gap_compressed_identity_from_alignment(ref_pos, cigar[cigar_start_index:cigar_end_index], read_seq[read_offset:], chrom_seq, True)


def passes_read_filter(mapq: int, pos: int, cigar: Sequence[Cigar], read_seq: bytes, chrom_seq: bytes, min_sv_mapq: int = 5,
                       min_gap_compressed_identity: float = 0.97) -> bool:
    """src/score_sv/mod.rs:788-797 (the flag filter `filter_out_alignment_record` is the caller's)."""
    if mapq < min_sv_mapq:
        return False
    return get_gap_compressed_identity_from_alignment(pos, cigar, read_seq, chrom_seq, True) >= min_gap_compressed_identity


def get_alignment_closest_to_target_ref_pos(alignment_start_ref_pos: int, cigar: Sequence[Cigar], target_ref_pos: int,
                                            target_match_type: str) -> Optional[Tuple[int, int]]:
    """(read_pos, ref_pos) of the aligned base closest to target_ref_pos, optionally constrained to one side of it."""
    closest: Optional[Tuple[int, int]] = None
    ref_pos, read_pos = alignment_start_ref_pos, 0
    for c in cigar:
        op, n = c
        if op in "X=M":
            start_diff = target_ref_pos - ref_pos
            end_diff = start_diff - n
            if start_diff >= 0 and end_diff < 0:
                closest = (read_pos + start_diff, ref_pos + start_diff)
                break
            if closest is not None:
                low_diff = min(abs(start_diff), abs(end_diff))
                if target_ref_pos - closest[1] <= low_diff:
                    break
            if abs(start_diff) < abs(end_diff):   # past the target: start of this match block
                closest = (read_pos, ref_pos) if target_match_type in (ANY, TARGET_OR_HIGHER) else None
            else:                                 # before the target: end of this match block
                closest = (read_pos + n - 1, ref_pos + n - 1) if target_match_type in (ANY, TARGET_OR_LOWER) else None
        ref_pos += _ref_advance(c)
        read_pos += _read_advance(c, True)
    return closest


def translate_ref_range_to_hardclipped_read_range(pos: int, cigar: Sequence[Cigar], ref_range: Tuple[int, int]) -> Optional[Tuple[int, int]]:
    rs, re_ = ref_range
    read_start = read_end = None
    read_pos, ref_pos, read_pos_match_end = 0, pos, 0
    for c in cigar:
        op, n = c
        if op in "X=M":
            if read_start is None:
                if ref_pos <= rs < ref_pos + n:
                    read_start = read_pos + (rs - ref_pos)
                elif rs < ref_pos < re_:
                    read_start = read_pos
            if ref_pos <= re_ < ref_pos + n:
                read_end = read_pos + (re_ - ref_pos)
                break
            read_pos_match_end = read_pos + n
        elif op in "DN":
            if ref_pos <= rs < ref_pos + n:
                read_start = read_pos - 1
            if ref_pos <= re_ < ref_pos + n:
                read_end = read_pos
                break
        ref_pos += _ref_advance(c)
        read_pos += _read_advance(c, True)
    if read_start is not None and read_end is None and re_ >= ref_pos:
        read_end = read_pos_match_end
    return (read_start, read_end) if read_start is not None and read_end is not None else None


def get_alignment_ref_segments(ref_pos: int, cigar: Sequence[Cigar], min_gap_size: int) -> List[Tuple[int, int]]:
    out, begin = [], ref_pos
    for c in cigar:
        op, n = c
        if op in "DN" and n >= min_gap_size:
            if ref_pos - begin > 0:
                out.append((begin, ref_pos))
            begin = ref_pos + n
        ref_pos += _ref_advance(c)
    if ref_pos - begin > 0:
        out.append((begin, ref_pos))
    return out


@dataclass
class ReadFlankInfo:  # src/score_sv/mod.rs ReadFlankInfo
    range: Tuple[int, int]
    start_truncation: int
    end_truncation: int


def get_read_breakend_flank_range(pos: int, cigar: Sequence[Cigar], read_len: int, breakend_dir: str, breakend_range: Tuple[int, int],
                                  assembly_range: Optional[Tuple[int, int]] = None) -> Optional[ReadFlankInfo]:
    """src/score_sv/mod.rs:321-432.  breakend_range = the breakend's segment range (its size - 1 is the homology length);
    assembly_range: register the read at the edge of the assembly region instead of at the breakend (the second,
    "AsmRegionEdge" registration)."""
    homology_len = breakend_range[1] - breakend_range[0] - 1
    assert homology_len >= 0

    def flank(start, end, start_trunc, end_trunc):
        return None if end - start < MIN_READ_SUPPORT_FLANK_SIZE else ReadFlankInfo((start, end), start_trunc, end_trunc)

    if breakend_dir == LEFT_ANCHOR:
        ref_target_pos = breakend_range[0]
        if assembly_range is not None:
            if assembly_range[0] >= ref_target_pos - REF_TO_READ_ANCHOR_OFFSET:
                return None
            anchor = assembly_range[0]
        else:
            anchor = ref_target_pos
        m = get_alignment_closest_to_target_ref_pos(pos, cigar, anchor - REF_TO_READ_ANCHOR_OFFSET, TARGET_OR_LOWER)
        if m is None:
            return None
        shifted_read_pos = m[0] + (ref_target_pos - m[1])
        if shifted_read_pos + 1 < READ_SUPPORT_ANCHOR_MIN_EDGE_DISTANCE:
            return None
        flank_start = shifted_read_pos + 1 + homology_len
        full_end = flank_start + READ_SUPPORT_FLANK_SIZE
        flank_end, end_trunc = (read_len, full_end - read_len) if full_end > read_len else (full_end, 0)
        return flank(flank_start, flank_end, 0, end_trunc)
    ref_target_pos = breakend_range[1]
    if assembly_range is not None:
        if assembly_range[1] <= ref_target_pos + REF_TO_READ_ANCHOR_OFFSET:
            return None
        anchor = assembly_range[1]
    else:
        anchor = ref_target_pos
    m = get_alignment_closest_to_target_ref_pos(pos, cigar, anchor + REF_TO_READ_ANCHOR_OFFSET, TARGET_OR_HIGHER)
    if m is None:
        return None
    shifted_read_pos = m[0] + (ref_target_pos - m[1])
    if read_len - shifted_read_pos < READ_SUPPORT_ANCHOR_MIN_EDGE_DISTANCE:
        return None
    flank_end = shifted_read_pos - homology_len
    full_start = flank_end - READ_SUPPORT_FLANK_SIZE
    flank_start, start_trunc = (0, -full_start) if full_start < 0 else (full_start, 0)
    return flank(flank_start, flank_end, start_trunc, 0)


# ------------------------------------------------------------------------------------------------ allele (pattern) side
def convert_breakend_to_ref_flank_range(breakend_dir: str, breakend_range: Tuple[int, int], chrom_len: int) -> Optional[Tuple[int, int]]:
    """convert_breakend_to_ref_flank (src/score_sv/mod.rs:36-78): the reference range the ref-allele flank is cut from.
    Left anchor: the 500 bases after the breakend segment; right anchor: the 500 bases before its start plus that base
    (GenomeSegment::asymmetric_expand_by clips at the chromosome ends)."""
    if breakend_dir == LEFT_ANCHOR:
        start = breakend_range[1]
        end = min(breakend_range[1] + READ_SUPPORT_FLANK_SIZE, chrom_len)
    else:
        end = min(breakend_range[0] + 1, chrom_len)
        start = max(breakend_range[0] - READ_SUPPORT_FLANK_SIZE, 0)
    if start < 0 or end <= start or end - start < MIN_READ_SUPPORT_FLANK_SIZE:
        return None
    return start, end


def convert_breakend_to_haplotype_contig_flank_range(contig_len: int, ref_offset: int, cigar: Sequence[Cigar], breakend_dir: str,
                                                     breakend_range: Tuple[int, int],
                                                     assembly_range: Optional[Tuple[int, int]] = None) -> Optional[Tuple[int, int]]:
    """convert_breakend_to_haplotype_contig_flank_seq (:106-222): the contig range the alt-allele flank is cut from; the
    contig alignment plays the role the read alignment plays in get_read_breakend_flank_range, without the 10-base anchor
    offset and without truncation (a flank that would leave the contig is dropped)."""
    homology_len = breakend_range[1] - breakend_range[0] - 1

    def flank(start, end):
        if start < 0 or end >= contig_len or end - start < MIN_READ_SUPPORT_FLANK_SIZE:
            return None
        return start, end

    if breakend_dir == LEFT_ANCHOR:
        ref_target_pos = breakend_range[0]
        anchor = assembly_range[0] if assembly_range is not None else ref_target_pos
        m = get_alignment_closest_to_target_ref_pos(ref_offset, cigar, anchor, TARGET_OR_LOWER)
        if m is None:
            return None
        flank_start = m[0] + (ref_target_pos - m[1]) + 1 + homology_len
        return flank(flank_start, flank_start + READ_SUPPORT_FLANK_SIZE)
    ref_target_pos = breakend_range[1]
    anchor = assembly_range[1] if assembly_range is not None else ref_target_pos
    m = get_alignment_closest_to_target_ref_pos(ref_offset, cigar, anchor, TARGET_OR_HIGHER)
    if m is None:
        return None
    flank_end = m[0] + (ref_target_pos - m[1]) - homology_len
    return flank(flank_end - READ_SUPPORT_FLANK_SIZE, flank_end)
