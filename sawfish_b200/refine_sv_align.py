"""Batched contig-to-reference alignment for refine_sv and the first CIGAR consumer.

Mirrors, around the alignment call, the reference's src/refine_sv/align/mod.rs:
  align_single_ref_region_assemblies :373-451  -> align_contigs_batch (one batch for every contig of a chunk of
                                                  clusters instead of new_large_indel_aligner + align per contig)
  get_single_region_refined_sv_candidates :260-365 -> get_single_region_indel_candidates (the CIGAR walk:
        Ins/Del >= min_indel_size inside the contig's high-quality range, kept only when flanked on both sides by
        an '=' run >= min_assembly_edge_anchor; large-insertion mode keeps insertions > 3000 only)
Not mirrored: create_indel_refined_sv_candidate's breakend homology / left-shift normalisation
(lib/rust-vc-utils/src/indel_breakend_homology.rs) and the 2-region alt-hap transform (segment_alignment.rs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .batch import LARGE_INDEL_PENALTIES, PAIR_DTYPE, PAIR_REVERSE, sawfish_clip, translate_rle
from .simple_alignment import SimpleAlignment, cigar_from_u32

MIN_INDEL_SIZE = 35            # --min-indel-size default, src/cli/discover.rs:131-132
MIN_ASSEMBLY_EDGE_ANCHOR = 50  # src/refine_sv/mod.rs:667
MIN_LARGE_INSERTION_LEN = 3000  # src/refine_sv/align/mod.rs:353-361


def align_contigs_batch(ctx, targets: Sequence[bytes], contigs: Sequence[Tuple[int, bytes]]):
    """targets: reference windows; contigs: (target index, contig sequence).
    Returns [(score, Optional[SimpleAlignment])] in contig order == PairwiseAligner::align per contig."""
    chunks, rows, off = [], [], 0
    t_offs = []
    for t in targets:
        a = np.frombuffer(bytes(t), np.uint8)
        assert len(a) > 0       # set_text: assert!(!text.is_empty())
        chunks.append(a)
        t_offs.append(off)
        off += len(a)
    for ti, c in contigs:
        a = np.frombuffer(bytes(c), np.uint8)
        assert len(a) > 0       # align: assert!(!pattern.is_empty())
        chunks.append(a)
        clip = sawfish_clip(len(a), len(targets[ti]))
        rows.append((off, t_offs[ti], len(a), len(targets[ti]), clip, clip, clip, clip, PAIR_REVERSE))
        off += len(a)
    arena = np.concatenate(chunks)
    pairs = np.array(rows, dtype=PAIR_DTYPE)
    scores, status, ops, ops_off = ctx.align_batch(LARGE_INDEL_PENALTIES, arena, pairs, True)
    assert (status == 0).all()  # assert_eq!(status, StatusAlgCompleted)
    out = []
    for i in range(len(pairs)):
        roff, cig, has_match = translate_rle(ops[int(ops_off[i]):int(ops_off[i + 1])])
        out.append((int(scores[i]), SimpleAlignment(roff, cigar_from_u32(cig)) if has_match else None))
    return out


@dataclass
class IndelCandidate:
    ref_pos: int        # chrom_pos at the start of the event
    contig_pos: int
    del_len: int
    ins_len: int
    cigar_segment_index: int


def get_single_region_indel_candidates(alignment: SimpleAlignment, chrom_pos: int, high_quality_range: Tuple[int, int],
                                       min_indel_size: int = MIN_INDEL_SIZE,
                                       min_assembly_edge_anchor: int = MIN_ASSEMBLY_EDGE_ANCHOR,
                                       is_large_insertion: bool = False) -> List[IndelCandidate]:
    """The CIGAR walk of get_single_region_refined_sv_candidates (src/refine_sv/align/mod.rs:277-363).
    chrom_pos: reference position of the first aligned base (chrom_segment_start + alignment.ref_offset).
    high_quality_range: [start, end) of the contig's high-quality range (IntRange::intersect_pos)."""
    contig_pos = 0
    found_anchor = False
    tmp: List[IndelCandidate] = []
    refined: List[IndelCandidate] = []
    for idx, (op, n) in enumerate(alignment.cigar):
        in_hq = high_quality_range[0] <= contig_pos < high_quality_range[1]
        if op == "=":
            if n >= min_assembly_edge_anchor:
                if found_anchor:
                    refined.extend(tmp)
                tmp = []
                found_anchor = True
        elif op == "D":
            if n >= min_indel_size and in_hq:
                tmp.append(IndelCandidate(chrom_pos, contig_pos, n, 0, idx))
        elif op == "I":
            if n >= min_indel_size and in_hq:
                tmp.append(IndelCandidate(chrom_pos, contig_pos, 0, n, idx))
        # update_ref_and_read_pos(cigar_segment, &mut chrom_pos, &mut contig_pos, ignore_hard_clip = true)
        if op in "MX=DN":
            chrom_pos += n
        if op in "MX=IS":
            contig_pos += n
    if is_large_insertion:
        return [c for c in refined if c.ins_len > MIN_LARGE_INSERTION_LEN]
    return refined


@dataclass
class NormalizedIndel:
    """Breakend description of one indel candidate after homology normalisation (the part of the reference's RefinedSV
    that depends on the alignment): half-open chromosome ranges, left-shifted positions, homology and insert sequence."""
    breakend1_range: Tuple[int, int]
    breakend2_range: Tuple[int, int]
    chrom_pos: int
    contig_pos: int
    breakend1_homology_seq: bytes
    insert_seq: bytes
    contig_pos_before_breakend1: int


def normalize_indel_candidate(chrom_pos: int, del_len: int, ins_len: int, contig_pos: int, contig_seq: bytes,
                              chrom_segment_start: int, chrom_segment_seq: bytes) -> NormalizedIndel:
    """Homology range of the indel, left shift, and the two breakend intervals: the alignment-dependent core of
    create_indel_refined_sv_candidate (src/refine_sv/align/mod.rs:71-150). This is why the placement of a clean indel
    inside its homology range by the aligner's tie-break is not observable in the candidate VCF."""
    from .indel_breakend_homology import get_indel_breakend_homology_info
    assert chrom_segment_start <= chrom_pos
    seg_pos = chrom_pos - chrom_segment_start
    (h0, h1), hom_seq = get_indel_breakend_homology_info(chrom_segment_seq, (seg_pos, seg_pos + del_len), contig_seq,
                                                         (contig_pos, contig_pos + ins_len))
    chrom_pos += h0  # left shift
    contig_pos += h0
    hom_len = h1 - h0
    b1 = (chrom_pos - 1, chrom_pos + hom_len)
    b2 = (chrom_pos - 1 + del_len, chrom_pos + del_len + hom_len)
    assert contig_pos >= 1
    return NormalizedIndel(b1, b2, chrom_pos, contig_pos, hom_seq, bytes(contig_seq[contig_pos:contig_pos + ins_len]),
                           contig_pos - 1)
