"""Two-region (breakpoint) alt-haplotype text of refine_sv, and the speculative batch of its contig alignments.

Mirror of reference src/refine_sv/align/mod.rs:
  construct_alt_hap_seq                 :708-737   segment1 + 100 x 'N' + segment2 (segment 2 reverse-complemented and / or first)
  get_middle_flank_size                 :755-770
  get_left_right_ref_flank_size         :776-793   interior flank capped at 2000
  get_chrom_info_from_target_segment    :46-64     (GenomeSegment::asymmetric_expand_by, lib/rust-vc-utils/src/genome_segment.rs:55-69)
  get_remote_breakend_info              :822-956   neighbouring breakpoint on the same haplotype: shortened flank + remote sequence
  extract_reference_segment_info        :974-1101
  get_two_region_alt_hap_info           :1125-1330
and of the alignment loop of align_multi_ref_region_assemblies (:1975-2199) in TWO-PHASE form: the reference aligns the
1500-flank contig of an assembly and only on failure the 2500-flank one (`break` at :2176); whether a contig fails is
decided by code that consumes the alignment, so every contig of every assembly of every cluster goes into ONE batch
(speculative: the 2500-flank alignments of assemblies whose first contig passes are computed and dropped), and the
sequential selection runs afterwards against the precomputed alignments.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .batch import PAIR_DTYPE, PAIR_REVERSE, sawfish_clip, translate_rle
from .contig_bam import rev_comp
from .simple_alignment import SimpleAlignment, cigar_from_u32

LEFT_ANCHOR, RIGHT_ANCHOR = 0, 1          # BreakendDirection
HAP_ALIGNMENT_REF_FLANK_SIZE = 3500       # RefineSVSettings::new, src/refine_sv/mod.rs:663
TWO_REGION_SPACER_SIZE = 100              # :673
MAX_INTERIOR_FLANK_SIZE = 2000            # align/mod.rs:785


@dataclass
class GenomeSegment:
    chrom_index: int
    start: int
    end: int

    def size(self) -> int:
        return self.end - self.start

    def asymmetric_expand_by(self, chrom_sizes: Sequence[int], left: int, right: int) -> Tuple[int, int]:
        new_start = max(self.start - left, 0)
        new_end = min(self.end + right, chrom_sizes[self.chrom_index])
        shifts = (self.start - new_start, new_end - self.end)
        self.start, self.end = new_start, new_end
        return shifts


@dataclass
class Breakend:
    segment: GenomeSegment
    dir: int


@dataclass
class Breakpoint:
    breakend1: Breakend
    breakend2: Optional[Breakend]

    def same_orientation(self) -> bool:       # src/breakpoint.rs:242-247
        return self.breakend2 is not None and self.breakend1.dir == self.breakend2.dir

    def get_breakend(self, index: int) -> Breakend:
        return self.breakend1 if index == 0 else self.breakend2


@dataclass
class FullBreakendNeighborInfo:
    breakend_index: int        # breakend_neighbor.breakend_index: which breakend of `breakpoint` is the local neighbour
    breakpoint: Breakpoint


@dataclass
class TwoRegionAltHapInfo:
    ref_segment1: GenomeSegment
    ref_segment2: GenomeSegment
    segment1_neighbor_extension_size: int
    segment2_neighbor_extension_size: int
    ref_segment2_revcomp: bool
    ref_segment2_after_ref_segment1: bool
    spacer_range: Tuple[int, int]
    left_boundary_pin: bool
    right_boundary_pin: bool
    alt_hap_seq: bytes
    extended_alt_hap_seq: bytes


def construct_alt_hap_seq(seg1: bytes, seg2: bytes, spacer_size: int, seg2_revcomp: bool, seg2_after_seg1: bool):
    if seg2_revcomp:
        seg2 = rev_comp(seg2)
    a, b = (seg1, seg2) if seg2_after_seg1 else (seg2, seg1)
    return a + b"N" * spacer_size + b, (len(a), len(a) + spacer_size)


def get_middle_flank_size(max_ref_flank_size: int, target_segments: Sequence[GenomeSegment], bp: Breakpoint) -> int:
    if (not bp.same_orientation()) and target_segments[0].chrom_index == target_segments[1].chrom_index:
        return min(max_ref_flank_size, max(0, target_segments[1].start - target_segments[0].end) // 2)
    return max_ref_flank_size


def get_left_right_ref_flank_size(max_ref_flank_size: int, is_segment_revcomped: bool, is_segment_first: bool) -> Tuple[int, int]:
    interior = min(MAX_INTERIOR_FLANK_SIZE, max_ref_flank_size)
    return (max_ref_flank_size, interior) if (is_segment_revcomped != is_segment_first) else (interior, max_ref_flank_size)


def get_chrom_info_from_target_segment(left: int, right: int, reference: Dict[int, bytes], chrom_sizes: Sequence[int],
                                       target_segment: GenomeSegment):
    seg = GenomeSegment(target_segment.chrom_index, target_segment.start, target_segment.end)
    al, ar = seg.asymmetric_expand_by(chrom_sizes, left, right)
    return seg, reference[seg.chrom_index][seg.start:seg.end], al, ar


def get_remote_breakend_info(full_ref_flank_size: int, reference, chrom_sizes, target_segment: GenomeSegment, target_breakend: Breakend,
                             neighbor: Optional[FullBreakendNeighborInfo]):
    """:822-956.  Returns None or (left_ref_flank_size, right_ref_flank_size, remote_seq)."""
    if neighbor is None:
        return None
    left_ref, right_ref = full_ref_flank_size, full_ref_flank_size
    local = neighbor.breakpoint.get_breakend(neighbor.breakend_index)
    remote = neighbor.breakpoint.get_breakend((neighbor.breakend_index + 1) % 2)
    assert local.dir != target_breakend.dir
    assert local.segment.chrom_index == target_breakend.segment.chrom_index
    left_remote = right_remote = 0
    if target_breakend.dir == LEFT_ANCHOR:
        max_left = target_segment.start - local.segment.start
        if max_left <= 0:
            return None
        left_ref = min(left_ref, max_left)
        left_remote = full_ref_flank_size - left_ref
    else:
        max_right = local.segment.end - target_segment.end
        if max_right <= 0:
            return None
        right_ref = min(right_ref, max_right)
        right_remote = full_ref_flank_size - right_ref
    assert left_remote >= 0 and right_remote >= 0 and (left_remote == 0 or right_remote == 0)
    if left_remote <= 0 and right_remote <= 0:
        return None
    is_remote_reversed = local.dir == remote.dir
    if is_remote_reversed:
        left_remote, right_remote = right_remote, left_remote
    if remote.dir == LEFT_ANCHOR:
        assert left_remote > 0
    else:
        assert right_remote > 0
    rseg = GenomeSegment(remote.segment.chrom_index, remote.segment.start, remote.segment.end)
    if left_remote > 0:
        rseg.end = rseg.start
    else:
        rseg.start = rseg.end
    _, remote_seq, _, _ = get_chrom_info_from_target_segment(left_remote, right_remote, reference, chrom_sizes, rseg)
    if is_remote_reversed:
        remote_seq = rev_comp(remote_seq)
    return left_ref, right_ref, remote_seq


def extract_reference_segment_info(full_ref_flank_size: int, reference, chrom_sizes, target_segment: GenomeSegment,
                                   target_breakend: Breakend, left_ref_flank_size: int, right_ref_flank_size: int,
                                   neighbor: Optional[FullBreakendNeighborInfo]):
    """:974-1101.  Returns (ref_segment, ref_segment_seq, extended_ref_segment_seq, actual_left_range, actual_right_range)."""
    remote = get_remote_breakend_info(full_ref_flank_size, reference, chrom_sizes, target_segment, target_breakend, neighbor)
    lf, rf = (remote[0], remote[1]) if remote else (full_ref_flank_size, full_ref_flank_size)
    ext_seg, ext_seq, al_range, ar_range = get_chrom_info_from_target_segment(lf, rf, reference, chrom_sizes, target_segment)
    al_seq, ar_seq = al_range, ar_range
    if remote:
        rs = remote[2]
        if target_breakend.dir == LEFT_ANCHOR:
            ext_seq = rs + ext_seq
            al_seq += len(rs)
        else:
            ext_seq = ext_seq + rs
            ar_seq += len(rs)
    left_seq_trim, right_seq_trim = max(al_seq - left_ref_flank_size, 0), max(ar_seq - right_ref_flank_size, 0)
    left_range_trim, right_range_trim = max(al_range - left_ref_flank_size, 0), max(ar_range - right_ref_flank_size, 0)
    assert left_seq_trim + right_seq_trim <= len(ext_seq)
    assert left_range_trim + right_range_trim <= ext_seg.size()
    seg = GenomeSegment(ext_seg.chrom_index, ext_seg.start + left_range_trim, ext_seg.end - right_range_trim)
    return seg, ext_seq[left_seq_trim:len(ext_seq) - right_seq_trim], ext_seq, al_range, ar_range


def get_two_region_alt_hap_info(reference: Dict[int, bytes], chrom_sizes: Sequence[int], target_segments: Sequence[GenomeSegment],
                                bp: Breakpoint, neighbors: Tuple[Optional[FullBreakendNeighborInfo], Optional[FullBreakendNeighborInfo]] = (None, None),
                                ref_flank_size: int = HAP_ALIGNMENT_REF_FLANK_SIZE, spacer_size: int = TWO_REGION_SPACER_SIZE) -> TwoRegionAltHapInfo:
    """:1125-1330."""
    be1, be2 = bp.breakend1, bp.breakend2
    assert be2 is not None and len(target_segments) == 2
    seg2_revcomp = be1.dir == be2.dir
    seg2_after_seg1 = be1.dir == LEFT_ANCHOR
    middle = get_middle_flank_size(ref_flank_size, target_segments, bp)
    l1, r1 = get_left_right_ref_flank_size(ref_flank_size, False, seg2_after_seg1)
    r1 = min(r1, middle)
    seg1, seq1, ext1, al1, ar1 = extract_reference_segment_info(ref_flank_size, reference, chrom_sizes, target_segments[0], be1, l1, r1, neighbors[0])
    l2, r2 = get_left_right_ref_flank_size(ref_flank_size, seg2_revcomp, not seg2_after_seg1)
    l2 = min(l2, middle)
    seg2, seq2, ext2, al2, ar2 = extract_reference_segment_info(ref_flank_size, reference, chrom_sizes, target_segments[1], be2, l2, r2, neighbors[1])
    assert len(seq1) >= seg1.size() and len(seq2) >= seg2.size()      # get_segment_neighbor_extension_size panics otherwise
    alt, spacer = construct_alt_hap_seq(seq1, seq2, spacer_size, seg2_revcomp, seg2_after_seg1)
    ext_alt, _ = construct_alt_hap_seq(ext1, ext2, spacer_size, seg2_revcomp, seg2_after_seg1)
    lp1, lp2, rp1, rp2 = al1 == 0, al2 == 0, ar1 == 0, ar2 == 0
    if seg2_after_seg1:
        pins = (rp1, rp2 if seg2_revcomp else lp2)
    else:
        pins = (lp2 if seg2_revcomp else rp2, lp1)
    return TwoRegionAltHapInfo(seg1, seg2, len(seq1) - seg1.size(), len(seq2) - seg2.size(), seg2_revcomp, seg2_after_seg1, spacer,
                               pins[0], pins[1], alt, ext_alt)


# ------------------------------------------------------------------------------------------------ speculative batch
@dataclass
class AssemblyContigs:
    """One assembly of a cluster: its contigs in the order the reference tries them (1500-flank first, then the
    2500-flank retry, src/refine_sv/mod.rs:664-666) and its read support."""
    contig_seqs: List[bytes]
    supporting_read_count: int = 0


class MultiRegionBatch:
    """Phase 1: add_cluster() per two-region cluster; phase 2: align(ctx) = ONE large-indel batch with CIGAR; phase 3:
    first_passing() replays the reference's contig loop of one assembly against the precomputed alignments."""

    def __init__(self):
        self.chunks: List[np.ndarray] = []
        self.off = 0
        self.rows: List[tuple] = []
        self.index: Dict[Tuple[int, int, int], int] = {}    # (cluster, assembly, contig) -> pair
        self.results: Optional[list] = None

    def _add(self, s: bytes) -> int:
        o = self.off
        self.chunks.append(np.frombuffer(s, np.uint8))
        self.off += len(s)
        return o

    def add_cluster(self, cluster_index: int, alt_hap_info: TwoRegionAltHapInfo, assemblies: Sequence[AssemblyContigs],
                    min_assembly_read_support: int = 0):
        t_off = self._add(alt_hap_info.alt_hap_seq)          # aligner.set_text(&alt_hap_info.alt_hap_seq), :2003
        tl = len(alt_hap_info.alt_hap_seq)
        for ai, asm in enumerate(assemblies):
            if asm.supporting_read_count < min_assembly_read_support:      # :2011-2019
                continue
            for ci, seq in enumerate(asm.contig_seqs):
                c = sawfish_clip(len(seq), tl)
                self.index[(cluster_index, ai, ci)] = len(self.rows)
                self.rows.append((self._add(seq), t_off, len(seq), tl, c, c, c, c, PAIR_REVERSE))

    def finish(self):
        arena = np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.uint8)
        return arena, np.array(self.rows, dtype=PAIR_DTYPE)

    def align(self, ctx) -> None:
        from .batch import LARGE_INDEL_PENALTIES
        arena, pairs = self.finish()
        self.results = []
        if not len(pairs):
            return
        scores, status, ops, ops_off = ctx.align_batch(LARGE_INDEL_PENALTIES, arena, pairs, True)
        assert (status == 0).all()            # == assert_eq!(status, StatusAlgCompleted), src/wfa2_utils.rs:117
        for i in range(len(pairs)):
            roff, cig, has_match = translate_rle(ops[int(ops_off[i]):int(ops_off[i + 1])])
            self.results.append((int(scores[i]), SimpleAlignment(roff, cigar_from_u32(cig)) if has_match else None))

    def alignment(self, cluster_index: int, assembly_index: int, contig_index: int):
        """(score, Option<SimpleAlignment>) as aligner.align(&assembly_contig.seq) returns it (:2029)."""
        return self.results[self.index[(cluster_index, assembly_index, contig_index)]]

    def first_passing(self, cluster_index: int, assembly_index: int, n_contigs: int,
                      accept: Callable[[int, SimpleAlignment], bool]) -> Optional[Tuple[int, SimpleAlignment]]:
        """The contig loop of :2022-2179 for one assembly: the first contig whose alignment exists and passes the
        reference's post-alignment filters (`accept`) wins; later contigs are not considered (`break`, :2176)."""
        for ci in range(n_contigs):
            _score, al = self.alignment(cluster_index, assembly_index, ci)
            if al is not None and accept(ci, al):
                return ci, al
        return None
