"""Read side of score_sv's pair enumeration (sawfish_b200/read_flanks.py): the reference's own unit tests for these helpers
(src/bam_utils.rs:378-640) replayed, plus the flank-range rules of src/score_sv/mod.rs:321-432 on constructed reads."""
import math

from sawfish_b200 import read_flanks as RF
from sawfish_b200.simple_alignment import cigar_from_string as cg


def test_gap_compressed_identity_reference_vectors():
    ref = b"AACGTGTTAACCCCT"  # contig_gap_test of the reference's test_data/test_reference.fa (bam_utils.rs:383)
    cases = [
        (b"AACGTGTTAACCCCT", "15=", 0, 1.0),
        (b"AACGTGTTAACCCCT", "15M", 0, 1.0),
        (b"AAAAACGTGTTAACCCCT", "3S15M100H", 0, 1.0),
        (b"GTTAACCCCT", "10M", 5, 1.0),
        (b"GTCAACCGCT", "10M", 5, 0.8),
        (b"GTATTANNACCAT", "2S4M2I4M1D1M100H", 5, 7.0 / 11.0),
    ]
    for seq, cigar, pos, want in cases:
        assert RF.get_gap_compressed_identity_from_alignment(pos, cg(cigar), seq, ref, True) == want, (seq, cigar)
    gci = RF.get_gap_compressed_identity_from_cigar_segment_range(2, b"GTAATCTTAC", cg("4M2I4M"), b"ACGTACGTACGT", 2, 3)
    assert math.isclose(gci, 3.0 / 4.0, rel_tol=0, abs_tol=1e-15)
    assert RF.passes_read_filter(20, 5, cg("10M"), b"GTTAACCCCT", ref) and not RF.passes_read_filter(20, 5, cg("10M"), b"GTCAACCGCT", ref)
    assert not RF.passes_read_filter(4, 5, cg("10M"), b"GTTAACCCCT", ref)


def test_alignment_closest_to_target_ref_pos_reference_vectors():
    t = 6
    assert RF.get_alignment_closest_to_target_ref_pos(1, cg("10M"), t, RF.ANY) == (5, 6)
    c = cg("2M8S")
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.ANY) == (1, 2)
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.TARGET_OR_LOWER) == (1, 2)
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.TARGET_OR_HIGHER) is None
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.TARGET) is None
    c = cg("2M5D8M")
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.ANY) == (2, 8)
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.TARGET_OR_HIGHER) == (2, 8)
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.TARGET_OR_LOWER) is None
    assert RF.get_alignment_closest_to_target_ref_pos(1, c, t, RF.TARGET) is None


def test_translate_ref_range_reference_vectors():
    # SAM POS is 1-based: "chr1 2" -> pos 1
    r = (6, 8)
    assert RF.translate_ref_range_to_hardclipped_read_range(1, cg("5H10M"), r) == (5, 7)
    assert RF.translate_ref_range_to_hardclipped_read_range(1, cg("6M1I3M"), r) == (5, 8)
    assert RF.translate_ref_range_to_hardclipped_read_range(1, cg("3M4D7M"), r) == (2, 3)
    r = (4, 8)
    assert RF.translate_ref_range_to_hardclipped_read_range(199, cg("10M"), r) is None
    assert RF.translate_ref_range_to_hardclipped_read_range(6, cg("5S5M"), r) == (5, 7)
    assert RF.translate_ref_range_to_hardclipped_read_range(0, cg("6M4S"), r) == (4, 6)
    assert RF.translate_ref_range_to_hardclipped_read_range(0, cg("4M10D6M"), r) == (3, 4)


def test_alignment_ref_segments_reference_vector():
    assert RF.get_alignment_ref_segments(100, cg("10M2D10M10D10M"), 10) == [(100, 122), (132, 142)]


def test_read_breakend_flank_range_rules():
    """src/score_sv/mod.rs:321-432 on an ungapped 2 000-base read starting at reference position 10 000."""
    pos, cigar, L = 10_000, cg("2000M"), 2000
    # left-anchored breakend at ref 11 000 with 3 bases of homology: flank starts right after breakend + homology
    f = RF.get_read_breakend_flank_range(pos, cigar, L, RF.LEFT_ANCHOR, (11_000, 11_004))
    assert (f.range, f.start_truncation, f.end_truncation) == ((1004, 1504), 0, 0)
    # close to the read end: truncated flank, the missing length is reported
    f = RF.get_read_breakend_flank_range(pos, cigar, L, RF.LEFT_ANCHOR, (11_700, 11_701))
    assert (f.range, f.end_truncation) == ((1701, 2000), 201)
    # less than 100 bases left: no flank; anchor closer than 200 bases to the read start: no flank
    assert RF.get_read_breakend_flank_range(pos, cigar, L, RF.LEFT_ANCHOR, (11_950, 11_951)) is None
    assert RF.get_read_breakend_flank_range(pos, cigar, L, RF.LEFT_ANCHOR, (10_100, 10_101)) is None
    # right-anchored breakend: the flank ends before breakend - homology
    f = RF.get_read_breakend_flank_range(pos, cigar, L, RF.RIGHT_ANCHOR, (11_000, 11_004))
    assert (f.range, f.start_truncation, f.end_truncation) == ((501, 1001), 0, 0)
    f = RF.get_read_breakend_flank_range(pos, cigar, L, RF.RIGHT_ANCHOR, (10_300, 10_301))
    assert (f.range, f.start_truncation) == ((0, 301), 199)
    # the read has a deletion across the breakend: the anchor 10 bases before it still maps, the shift rule applies
    gapped = cg("995M30D1005M")
    f = RF.get_read_breakend_flank_range(pos, gapped, 2000, RF.LEFT_ANCHOR, (11_000, 11_001))
    assert f.range == (1001, 1501)   # anchor 10 990 -> read 990, shifted by the 10 reference bases to the breakend
    # assembly-region registration: anchored at the region edge, refused when the edge is within 10 bases of the breakend
    f = RF.get_read_breakend_flank_range(pos, cigar, L, RF.LEFT_ANCHOR, (11_000, 11_001), assembly_range=(10_900, 11_200))
    assert f.range == (1001, 1501)
    assert RF.get_read_breakend_flank_range(pos, cigar, L, RF.LEFT_ANCHOR, (11_000, 11_001), assembly_range=(10_995, 11_200)) is None
