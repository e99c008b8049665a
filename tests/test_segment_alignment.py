"""Two-region CIGAR consumer and breakend homology: the reference's own known answers.

Golden vectors: src/refine_sv/align/segment_alignment.rs:366-668 (six tests) and
lib/rust-vc-utils/src/indel_breakend_homology.rs:79-147 (two tests), restated as data.
"""
import numpy as np
import pytest

from sawfish_b200.indel_breakend_homology import get_indel_breakend_homology_info
from sawfish_b200.segment_alignment import (TwoRegionAltHapInfo, does_alignment_support_alt_hap,
                                            transform_alt_hap_alignment_into_left_right_components)
from sawfish_b200.simple_alignment import SimpleAlignment, cigar_from_string, cigar_to_string


def base_info(seq=b"TTTTACTGACTGTTTNNNNTTTACTGACTGTTTT", spacer=(15, 19), left_pin=False, right_pin=False):
    return TwoRegionAltHapInfo(alt_hap_seq=seq, spacer_range=spacer, left_boundary_pin=left_pin, right_boundary_pin=right_pin)


def run(cigar, info, contig, anchor=3, ref_offset=4):
    return transform_alt_hap_alignment_into_left_right_components(anchor, SimpleAlignment(ref_offset, cigar_from_string(cigar)),
                                                                  info, contig)


def check(res, left_off, left_cigar, right_off, right_cigar, hom_range, hom_seq=None):
    assert res is not None
    assert res.left_alignment.ref_offset == left_off and cigar_to_string(res.left_alignment.cigar) == left_cigar
    assert res.right_alignment.ref_offset == right_off and cigar_to_string(res.right_alignment.cigar) == right_cigar
    assert res.left_right_breakend_homology_range == hom_range
    if hom_seq is not None:
        assert res.left_right_breakend_homology_seq == hom_seq


def test_plain_split_and_anchor_requirement():  # segment_alignment.rs:389-430
    contig = b"ACTGACTGACTGACTG"
    assert run("8=10D8=", base_info(), contig, anchor=10) is None
    check(run("8=10D8=", base_info(), contig), 4, "8=8S", 3, "8S8=", (0, 0))


def test_breakend_homology():  # :432-470
    info = base_info(seq=b"TTTTACTGACTGACCNNNNTTTACTGACTGTTTT")
    check(run("8=10D8=", info, b"ACTGACTGACTGACTG"), 4, "8=8S", 3, "8S8=", (0, 2), b"AC")


def test_insertion_right_of_the_breakend():  # :472-504
    check(run("8=10D4I8=", base_info(), b"ACTGACTGCCCCACTGACTG"), 4, "8=12S", 3, "12S8=", (0, 0))


def test_insertion_left_of_the_breakend():  # :506-538
    check(run("8=4I10D8=", base_info(), b"ACTGACTGCCCCACTGACTG"), 4, "8=12S", 3, "12S8=", (0, 0))


def test_left_pin():  # :540-602
    seq = b"TTTTACTGACTGNNNNTTTACTGACTGTTTT"
    contig = b"ACTGACTGACTGACTG"
    check(run("8=7D8=", base_info(seq, (12, 16), left_pin=True), contig), 4, "8=8S", 3, "8S8=", (0, 0))
    assert run("8=7D8=", base_info(seq, (12, 16), left_pin=False), contig) is None


def test_right_pin():  # :604-668
    seq = b"TTTTACTGACTGTTTNNNNACTGACTGTTTT"
    check(run("8=7D8=", base_info(seq, (15, 19), right_pin=True), b"ACTGACTGACTGACTG"), 4, "8=8S", 0, "8S8=", (0, 0))


def test_rejections():
    info = base_info()
    contig = b"ACTGACTGACTGACTG"
    # a match block across the spacer, and an op that ends inside it (segment_alignment.rs:232-248)
    assert run("8=3X4=3X8=", info, b"A" * 26) is None
    assert run("8=5D3D8=", info, contig) is None
    assert not does_alignment_support_alt_hap(3, SimpleAlignment(4, cigar_from_string("8=10D2=")), 15, 19)


def test_indel_homology_range():  # indel_breakend_homology.rs:79-119
    seq1, seq2 = b"ABCDDABC", b"ABCDDDABC"
    assert get_indel_breakend_homology_info(seq2, (3, 4), seq1, (3, 3)) == ((0, 2), b"DD")
    assert get_indel_breakend_homology_info(seq1, (3, 3), seq2, (3, 4)) == ((0, 2), b"DD")
    assert get_indel_breakend_homology_info(seq2, (5, 6), seq1, (5, 5)) == ((-2, 0), b"DD")
    assert get_indel_breakend_homology_info(seq1, (5, 5), seq2, (5, 6)) == ((-2, 0), b"DD")


def test_indel_homology_range_edges():  # :121-147
    assert get_indel_breakend_homology_info(b"DDDDDDABC", (3, 4), b"DDDDABC", (2, 2)) == ((-2, 2), b"DDDD")
    assert get_indel_breakend_homology_info(b"ABCDDDDDD", (3, 4), b"ABCDDDD", (3, 3)) == ((0, 4), b"DDDD")


@pytest.mark.gpu
def test_two_region_pair_through_the_cuda_path():
    """Breakpoint-shaped pair (segment 1 + 100 x N + segment 2, src/refine_sv/align/mod.rs:712-737) aligned by K2,
    its CIGAR split by the consumer: the spacer must be crossed by one deletion and both sides anchored."""
    from sawfish_b200.batch import AlignContext
    from sawfish_b200.refine_sv_align import align_contigs_batch
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seg1, seg2 = acgt[rng.integers(0, 4, 4000)].tobytes(), acgt[rng.integers(0, 4, 4200)].tobytes()
    alt_hap = seg1 + b"N" * 100 + seg2
    ctx = AlignContext(0)
    info = TwoRegionAltHapInfo(alt_hap_seq=alt_hap, spacer_range=(4000, 4100))
    # breakends 50 bases inside each segment: the deletion spans the spacer, both sides anchored
    contig = seg1[2150:3950] + seg2[50:1850]
    (score, aln), = align_contigs_batch(ctx, [alt_hap], [(0, contig)])
    assert aln is not None and aln.ref_offset == 2150 and cigar_to_string(aln.cigar) == "1800=200D1800="
    res = transform_alt_hap_alignment_into_left_right_components(50, aln, info, contig)
    assert res is not None
    assert res.left_alignment.ref_offset == 2150 and cigar_to_string(res.left_alignment.cigar) == "1800=1800S"
    assert res.right_alignment.ref_offset == 50 and cigar_to_string(res.right_alignment.cigar) == "1800S1800="
    lo, hi = res.left_right_breakend_homology_range
    assert lo <= 0 <= hi and len(res.left_right_breakend_homology_seq) == hi - lo
    # contig ending exactly on the last base of segment 1: accepted only when that segment is pinned
    # (segment_alignment.rs:196-207)
    contig = seg1[-1800:] + seg2[:1800]
    (score, aln), = align_contigs_batch(ctx, [alt_hap], [(0, contig)])
    assert aln is not None and aln.ref_offset == 2200 and cigar_to_string(aln.cigar) == "1800=100D1800="
    assert transform_alt_hap_alignment_into_left_right_components(50, aln, info, contig) is None
    info.left_boundary_pin = True
    res = transform_alt_hap_alignment_into_left_right_components(50, aln, info, contig)
    assert res is not None and res.right_alignment.ref_offset == 0 and cigar_to_string(res.right_alignment.cigar) == "1800S1800="


def test_collapse_indels_at_breakpoint():  # src/refine_sv/align/mod.rs:2271-2341
    from sawfish_b200.segment_alignment import collapse_indels_at_breakpoint
    info = TwoRegionAltHapInfo(alt_hap_seq=b"", spacer_range=(300, 400))
    for cigar_in, expect in [("100=300D100=", "100=300D100="), ("87=10I3=300D100=", "87=13I303D100="),
                             ("100=300D3=10I87=", "100=13I303D87="), ("96=1X3=300D3=1X96=", "96=8I308D96=")]:
        aln = SimpleAlignment(100, cigar_from_string(cigar_in))
        collapse_indels_at_breakpoint(info, aln)
        assert cigar_to_string(aln.cigar) == expect and aln.ref_offset == 100


def test_normalize_indel_candidate_left_shifts_into_homology():
    from sawfish_b200.refine_sv_align import normalize_indel_candidate
    # ref ABCDDDABC vs contig ABCDDABC: a 1-base deletion reported at its right-most position (ref 5) moves to ref 3
    ref, contig = b"ABCDDDABC", b"ABCDDABC"
    n = normalize_indel_candidate(chrom_pos=1005, del_len=1, ins_len=0, contig_pos=5, contig_seq=contig,
                                  chrom_segment_start=1000, chrom_segment_seq=ref)
    assert (n.chrom_pos, n.contig_pos, n.breakend1_homology_seq, n.insert_seq) == (1003, 3, b"DD", b"")
    assert n.breakend1_range == (1002, 1005) and n.breakend2_range == (1003, 1006) and n.contig_pos_before_breakend1 == 2
    # the same event reported left-most is a fixed point
    m = normalize_indel_candidate(1003, 1, 0, 3, contig, 1000, ref)
    assert (m.chrom_pos, m.breakend1_range, m.breakend2_range) == (1003, (1002, 1005), (1003, 1006))
