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


# ------------------------------------------------------------------------------------------------ split reads
# For breakpoint (two-region) SVs a read reaches the locus as several BAM records; a record's flank is only used for a
# breakend when THAT record is the split segment next to the breakend and its neighbour in read order sits on the mate
# breakend (src/score_sv/mod.rs:484-640).
def get_read_clip_positions(cigar: Sequence[Cigar], ignore_hard_clip: bool) -> Tuple[int, int, int]:
    """(first position after the left clip, first position of the right clip, read length) —
    lib/rust-vc-utils/src/bam_utils/cigar/mod.rs:85-116."""
    read_pos = left_clip = right_clip = 0
    left = True
    for c in cigar:
        op, n = c
        if op == "S" or (op == "H" and not ignore_hard_clip):
            if left:
                left_clip += n
            else:
                right_clip += n
        elif op != "H":
            left = False
        read_pos += _read_advance(c, ignore_hard_clip)
    return left_clip, read_pos - right_clip, read_pos


@dataclass
class SeqOrderSplitReadSegment:  # lib/rust-vc-utils/src/bam_utils/split_read.rs:14-33
    seq_order_read_start: int
    seq_order_read_end: int
    chrom_index: int
    pos: int
    is_fwd_strand: bool
    cigar: List[Cigar]
    mapq: int
    from_primary_bam_record: bool


def get_seq_order_read_split_segments(label_to_index, tid: int, pos: int, is_reverse: bool, cigar: Sequence[Cigar], mapq: int,
                                      sa_tag: Optional[str]) -> List[SeqOrderSplitReadSegment]:
    """All alignment segments of a read (this record + its SA tag) ordered by their start in SEQUENCING order of the read
    (split_read.rs:56-161).  Hard clips count (ignore_hard_clip = false): supplementary records are hard-clipped."""
    from .simple_alignment import cigar_from_string

    def seq_order(read_start, read_end, read_size, fwd):
        return (read_start, read_end) if fwd else (read_size - read_end, read_size - read_start)

    rs, re_, size = get_read_clip_positions(cigar, False)
    s, e = seq_order(rs, re_, size, not is_reverse)
    segs = [SeqOrderSplitReadSegment(s, e, tid, pos, not is_reverse, list(cigar), mapq, True)]
    if sa_tag:
        for word in sa_tag.split(";"):
            if not word:
                continue
            rname, pos1, strand, cig, mq, _nm = word.split(",")
            c2 = cigar_from_string(cig)
            assert any(op in "M=X" for op, _ in c2), "Bam record split segment is unaligned"
            rs, re_, size2 = get_read_clip_positions(c2, False)
            assert size2 == size
            fwd = strand == "+"
            s, e = seq_order(rs, re_, size2, fwd)
            segs.append(SeqOrderSplitReadSegment(s, e, label_to_index[rname], int(pos1) - 1, fwd, c2, int(mq), False))
        segs.sort(key=lambda x: x.seq_order_read_start)  # stable, like sort_by_key
    for x in segs:
        assert x.seq_order_read_start < x.seq_order_read_end, "Can't parse consistent split read information from SA tag"
    return segs


@dataclass
class BreakendMatchSplitReadInfo:  # src/score_sv/mod.rs:480-483
    seq_order_read_split_segments: List[Tuple[int, int, int]]   # (chrom_index, start, end) per segment, read order
    primary_split_segment_index: int


def get_split_read_info(label_to_index, tid, pos, is_reverse, cigar, mapq, sa_tag, single_region_refinement: bool) -> Optional[BreakendMatchSplitReadInfo]:
    """:602-633.  None = no check needed (indel-like SV, or the read is not split)."""
    if single_region_refinement or not sa_tag:
        return None
    out, primary = [], 0
    for i, seg in enumerate(get_seq_order_read_split_segments(label_to_index, tid, pos, is_reverse, cigar, mapq, sa_tag)):
        if seg.from_primary_bam_record:
            primary = i
        out.append((seg.chrom_index, seg.pos, seg.pos + sum(n for op, n in seg.cigar if op in "MDN=X")))
    return BreakendMatchSplitReadInfo(out, primary)


def get_segment_breakend_distance(segment: Tuple[int, int, int], breakend_chrom: int, breakend_dir: str, breakend_start: int) -> Optional[int]:
    """:485-500: a left-anchored breakend is matched by a segment END, a right-anchored one by a segment START."""
    if segment[0] != breakend_chrom:
        return None
    return abs((segment[2] if breakend_dir == LEFT_ANCHOR else segment[1]) - breakend_start)


def is_segment_breakend_match(segment_index: int, breakend, info: BreakendMatchSplitReadInfo) -> bool:
    """:504-522: no other split segment of the read is closer to the breakend.  breakend = (chrom, dir, range start)."""
    d = [get_segment_breakend_distance(s, *breakend) for s in info.seq_order_read_split_segments]
    if d[segment_index] is None:
        return False
    return all(x >= d[segment_index] for x in d if x is not None)


def does_read_segment_match_breakend(is_reverse: bool, target_breakend, mate_breakend, info: BreakendMatchSplitReadInfo) -> bool:
    """:526-565: this record is the closest segment to the target breakend AND the adjacent segment in read order (the next
    one for a left anchor on a forward record, the previous one otherwise) is the closest to the mate breakend."""
    t = info.primary_split_segment_index
    if not is_segment_breakend_match(t, target_breakend, info):
        return False
    is_plus_one = (target_breakend[1] == LEFT_ANCHOR) ^ is_reverse
    if is_plus_one:
        if t + 1 >= len(info.seq_order_read_split_segments):
            return False
        mate_index = t + 1
    else:
        if t == 0:
            return False
        mate_index = t - 1
    return is_segment_breakend_match(mate_index, mate_breakend, info)
