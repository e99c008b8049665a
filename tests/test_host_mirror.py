"""Host-side mirrors of the reference's result type, following the reference's own unit tests."""
from sawfish_b200.simple_alignment import (SimpleAlignment, cigar_clip_alignment_ref_edges, cigar_from_string,
                                           cigar_to_string, clip_alignment_ref_edges, compress_cigar,
                                           cigar_from_u32, cigar_to_u32)
from sawfish_b200.batch import sawfish_clip


def test_get_reverse():
    # reference src/simple_alignment.rs:78-109
    al = SimpleAlignment(1, [("M", 1)]).get_reverse(4)
    assert al.ref_offset == 2 and al.cigar == [("M", 1)]
    al = SimpleAlignment(1, [("M", 2)]).get_reverse(4)
    assert al.ref_offset == 1 and al.cigar == [("M", 2)]
    al = SimpleAlignment(2, [("M", 2), ("D", 2), ("X", 1)]).get_reverse(20)
    assert al.ref_offset == 13 and al.cigar == [("X", 1), ("D", 2), ("M", 2)]


def test_clip_alignment_ref_edges():
    # reference lib/rust-vc-utils/src/bam_utils/cigar/clip_alignment.rs:185-211
    out, shift = cigar_clip_alignment_ref_edges([("S", 3), ("M", 15)], 5, 2)
    assert out == [("S", 8), ("M", 8), ("S", 2)] and shift == 5
    out, shift = cigar_clip_alignment_ref_edges([("S", 3), ("M", 2), ("D", 3), ("M", 13)], 5, 2)
    assert out == [("S", 5), ("M", 11), ("S", 2)] and shift == 5
    al = clip_alignment_ref_edges(SimpleAlignment(2, [("=", 20)]), 30, 5, 10)
    assert al.ref_offset == 5 and cigar_to_string(al.cigar) == "3S15=2S"


def test_cigar_codecs():
    s = "4S3=2X4=3D4=1I"
    c = cigar_from_string(s)
    assert cigar_to_string(c) == s and cigar_from_u32(cigar_to_u32(c)) == c
    assert compress_cigar([("M", 1), ("M", 2), ("I", 0), ("D", 3)]) == [("M", 3), ("D", 3)]


def test_sawfish_clip_truncation():
    # f64 -> i32 truncation of 0.4 * len (reference src/wfa2_utils.rs:111-115)
    assert sawfish_clip(6, 10) == 2 and sawfish_clip(500, 500) == 200 and sawfish_clip(7, 1000) == 2
    assert sawfish_clip(1, 1) == 0


def test_prob_utils_reference_vectors():
    """The reference's unit tests of the phred / normalisation helpers the genotype chain uses (src/prob_utils.rs:72-93)."""
    import math
    from sawfish_b200 import score_sv as S
    assert S._rust_round(S.error_prob_to_phred(0.001)) == 30                # test_error_prob_to_qphred
    assert S.ln_error_prob_to_qphred(math.log(0.001)) == 30                 # test_ln_error_prob_to_qphred
    x = [math.log(v) for v in (0.001, 0.001, 0.002, 0.001)]
    assert S.normalize_ln_distro(x) == 2                                    # test_normalize_ln_distro
    assert math.isclose(x[0], 0.2, rel_tol=0, abs_tol=4 * 2.0 ** -52) and math.isclose(x[2], 0.4, rel_tol=0, abs_tol=4 * 2.0 ** -52)
