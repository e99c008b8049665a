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


# ---------------------------------------------------------------------------------------------------------------------
# One locus end to end from ALIGNMENTS (not pre-cut flanks): reference, an assembled alt contig with its alignment, reads
# of both haplotypes with theirs -> read filter -> read / allele flank ranges -> pair batch -> scores -> AD -> genotype.
import numpy as np  # noqa: E402
import pytest  # noqa: E402

from sawfish_b200 import score_sv as S  # noqa: E402
from sawfish_b200 import synth  # noqa: E402


def _locus_pairs(seed, gt):
    rng = np.random.default_rng(seed)
    ref = synth.random_seq(rng, 8000)
    d0, dl = 4000, 300                                            # deletion of ref[4000:4300)
    contig = np.concatenate([ref[2500:d0], ref[d0 + dl:5800]])    # assembled alt haplotype
    contig_aln = (2500, [("=", 1500), ("D", dl), ("=", 1500)])
    be1 = (RF.LEFT_ANCHOR, (d0 - 1, d0))                          # VCF POS = base before the deletion
    be2 = (RF.RIGHT_ANCHOR, (d0 + dl - 1, d0 + dl))               # END - 1, see candidate_sv.process_record_to_anno_refined_sv
    asm = (3900, 4400)
    b = S.PairBatchBuilder()
    truth = {}
    for ri in range(30):
        from_alt = rng.random() < (0.0, 0.5, 1.0)[gt]
        start = int(rng.integers(2600, 3300))
        if from_alt:
            hap = np.concatenate([ref[start:d0], ref[d0 + dl:start + 2200]])
            cigar = [("=", d0 - start), ("D", dl), ("=", len(hap) - (d0 - start))]
        else:
            hap = ref[start:start + 2200]
            cigar = [("=", len(hap))]
        read = hap.copy()                                          # substitutions only: the alignment stays valid
        for p in rng.integers(0, len(read), 2):
            read[int(p)] = (read[int(p)] + 1) if read[int(p)] != ord("T") else ord("A")
        read = np.frombuffer(bytes(bytearray(read.tolist())).translate(bytes.maketrans(b"BDHU", b"CGTA")), np.uint8)
        mcig = [("M", n) if op == "=" else (op, n) for op, n in cigar]
        if not RF.passes_read_filter(60, start, mcig, read.tobytes(), ref.tobytes()):
            continue
        qname = f"read{ri:03d}"
        truth[qname] = from_alt
        for bi, (bdir, brange) in enumerate((be1, be2)):
            ref_rng = RF.convert_breakend_to_ref_flank_range(bdir, brange, len(ref))
            alt_std = RF.convert_breakend_to_haplotype_contig_flank_range(len(contig), contig_aln[0], contig_aln[1], bdir, brange)
            alt_ext = RF.convert_breakend_to_haplotype_contig_flank_range(len(contig), contig_aln[0], contig_aln[1], bdir, brange, asm)
            for reg, arng, alt_rng in ((S.STANDARD, None, alt_std), (S.ASM_REGION_EDGE, asm, alt_ext)):
                f = RF.get_read_breakend_flank_range(start, mcig, len(read), bdir, brange, arng)
                if f is None:
                    continue
                # the ref allele has no second registration: its standard flank is scored against both read registrations
                # (ref_flank_seqs.breakend_flanks[breakend_index][standard_registration_index], src/score_sv/mod.rs:842-846)
                ref_seq = ref[ref_rng[0]:ref_rng[1]] if ref_rng is not None else None
                alt_seq = contig[alt_rng[0]:alt_rng[1]] if alt_rng is not None else None
                b.add_read_flank(0, qname, bi, reg, S.ReadFlank(read[f.range[0]:f.range[1]], f.start_truncation, f.end_truncation),
                                 ref_seq, alt_seq, [])
    return b, truth


def _genotype(b, align_fn):
    arena, pairs = b.finish()
    scores = align_fn(arena, pairs)
    locus = S.scatter_scores(b.slots, scores, {0: []})
    ad = S.set_allele_depth_counts(locus.get(0, {}), sv_haplotype_index=0, overlapping_haplotype_index=None,
                                   is_alt_on_top_haplotype_rank=True, on_multiple_sample_haplotypes=False, read_to_hap={})
    samples, qual = S.get_sv_qualities_from_allele_counts([ad])
    return ad, samples[0], qual, len(pairs), scores


def _oracle_fn():
    from oracle import oracle as O

    def f(arena, pairs):
        sc, st, *_ = O.align_batch(O.CONSENSUS, arena, pairs.astype(O.PAIR_DTYPE), want_ops=True, n_threads=4)
        assert (st == 0).all()
        return sc
    return f


@pytest.mark.parametrize("gt", [0, 1, 2])
def test_locus_from_alignments_recovers_the_genotype(gt):
    b, truth = _locus_pairs(100 + gt, gt)
    ad, sample, qual, n_pairs, _ = _genotype(b, _oracle_fn())
    n_alt = sum(truth.values())
    assert n_pairs > 60
    assert ad[1] >= 0.8 * n_alt and ad[0] >= 0.8 * (len(truth) - n_alt)   # reads vote for the haplotype they came from
    assert sample.gt == (S.Genotype.Ref, S.Genotype.Het, S.Genotype.Hom)[gt]


@pytest.mark.gpu
def test_locus_from_alignments_gpu_equals_oracle():
    from sawfish_b200.batch import AlignContext
    ctx = AlignContext(0)
    b, _ = _locus_pairs(101, 1)
    want = _genotype(b, _oracle_fn())
    got = _genotype(b, S.gpu_align_fn(ctx))
    assert (got[4] == want[4]).all() and got[0] == want[0] and got[2] == want[2]
    ctx.close()


def test_seq_order_read_split_segments_reference_vectors():
    """lib/rust-vc-utils/src/bam_utils/split_read.rs:202-237."""
    l2i = {"chr0": 0, "chr1": 1, "chr2": 2}

    def flat(segs):
        from sawfish_b200.simple_alignment import cigar_to_string
        return [(s.seq_order_read_start, s.seq_order_read_end, s.chrom_index, s.pos, s.is_fwd_strand, cigar_to_string(s.cigar),
                 s.from_primary_bam_record) for s in segs]

    got = RF.get_seq_order_read_split_segments(l2i, 2, 9, False, cg("10S5M5S"), 60, None)
    assert flat(got) == [(10, 15, 2, 9, True, "10S5M5S", True)]
    sa = "chr0,20,-,5M15S,60,0;chr0,100,+,5S5M10S,60,0;chr1,200,-,15S5M,60,0;"
    got = RF.get_seq_order_read_split_segments(l2i, 2, 9, False, cg("10S5M5S"), 60, sa)
    assert flat(got) == [(0, 5, 1, 199, False, "15S5M", False), (5, 10, 0, 99, True, "5S5M10S", False),
                         (10, 15, 2, 9, True, "10S5M5S", True), (15, 20, 0, 19, False, "5M15S", False)]
    assert all(s.mapq == 60 for s in got)


def test_split_read_segment_matches_its_breakend():
    """A translocation-like junction chr0:5000 (left anchor) -> chr1:9000 (right anchor) and a read split across it."""
    l2i = {"chr0": 0, "chr1": 1}
    be1, be2 = (0, RF.LEFT_ANCHOR, 4999), (1, RF.RIGHT_ANCHOR, 9000)
    # primary record = first 1 000 bases on chr0 ending at the junction; supplementary = last 1 000 bases on chr1
    info = RF.get_split_read_info(l2i, 0, 4000, False, cg("1000M1000S"), 60, "chr1,9001,+,1000S1000M,60,0;", False)
    assert info.primary_split_segment_index == 0 and info.seq_order_read_split_segments == [(0, 4000, 5000), (1, 9000, 10000)]
    assert RF.does_read_segment_match_breakend(False, be1, be2, info)         # this record carries breakend 1 ...
    assert not RF.does_read_segment_match_breakend(False, be2, be1, info)     # ... but not breakend 2
    # the same read seen from its supplementary record
    info2 = RF.get_split_read_info(l2i, 1, 9000, False, cg("1000H1000M"), 60, "chr0,4001,+,1000M1000S,60,0;", False)
    assert info2.primary_split_segment_index == 1
    assert RF.does_read_segment_match_breakend(False, be2, be1, info2) and not RF.does_read_segment_match_breakend(False, be1, be2, info2)
    # a third, closer segment elsewhere on chr0 steals breakend 1
    info3 = RF.get_split_read_info(l2i, 0, 3000, False, cg("500M2500S"), 60, "chr0,4501,+,500S500M2000S,60,0;chr1,9001,+,2000S1000M,60,0;", False)
    assert not RF.does_read_segment_match_breakend(False, be1, be2, info3)
    assert RF.get_split_read_info(l2i, 0, 4000, False, cg("2000M"), 60, None, False) is None       # not split: no check
    assert RF.get_split_read_info(l2i, 0, 4000, False, cg("1000M1000S"), 60, "chr1,9001,+,1000S1000M,60,0;", True) is None


def test_cigar_walkers_against_per_base_maps():
    """Independent check of the cigar walkers: expand random alignments into explicit read<->ref base maps and compare."""
    rng = np.random.default_rng(77)
    for _ in range(300):
        cigar, ref_pos0 = [], int(rng.integers(0, 50))
        if rng.random() < 0.4:
            cigar.append(("S", int(rng.integers(1, 8))))
        n_blocks = int(rng.integers(1, 5))
        for b in range(n_blocks):
            cigar.append((str(rng.choice(["M", "=", "X"])), int(rng.integers(3, 12))))
            if b + 1 < n_blocks:
                cigar.append((str(rng.choice(["I", "D"])), int(rng.integers(1, 6))))
        if rng.random() < 0.4:
            cigar.append(("S", int(rng.integers(1, 8))))
        # per-base map of aligned (read_pos, ref_pos)
        aligned, read_pos, ref_pos = [], 0, ref_pos0
        for op, n in cigar:
            if op in "M=X":
                aligned += [(read_pos + i, ref_pos + i) for i in range(n)]
            read_pos += n if op in "MIS=X" else 0
            ref_pos += n if op in "MDN=X" else 0
        ref_to_read = {r: q for q, r in aligned}
        lo, hi = aligned[0][1], aligned[-1][1]
        for target in range(lo - 3, hi + 4):
            got = RF.get_alignment_closest_to_target_ref_pos(ref_pos0, cigar, target, RF.ANY)
            if target in ref_to_read:                                   # an aligned base: exact answer in every mode
                assert got == (ref_to_read[target], target)
                assert RF.get_alignment_closest_to_target_ref_pos(ref_pos0, cigar, target, RF.TARGET) == got
            else:
                assert got is not None
                best = min(abs(r - target) for _q, r in aligned)
                assert abs(got[1] - target) == best and ref_to_read[got[1]] == got[0]
                lower = RF.get_alignment_closest_to_target_ref_pos(ref_pos0, cigar, target, RF.TARGET_OR_LOWER)
                higher = RF.get_alignment_closest_to_target_ref_pos(ref_pos0, cigar, target, RF.TARGET_OR_HIGHER)
                assert lower is None or lower[1] < target
                assert higher is None or higher[1] > target
        # a ref range whose two ends are aligned bases maps to exactly their read positions
        a, b = sorted(int(x) for x in rng.choice(len(aligned), 2, replace=False))
        rng_ref = (aligned[a][1], aligned[b][1])
        assert RF.translate_ref_range_to_hardclipped_read_range(ref_pos0, cigar, rng_ref) == (aligned[a][0], aligned[b][0])
        # gap-compressed identity: =/X counted as written, M compared base by base, one event per indel
        ref_seq = bytes(rng.choice(list(b"ACGT"), ref_pos + 5).tolist())
        read = bytearray(rng.choice(list(b"ACGT"), read_pos).tolist())
        match = mism = 0
        q, r = 0, ref_pos0
        for op, n in cigar:
            if op == "=":
                match += n
            elif op == "X":
                mism += n
            elif op == "M":
                for i in range(n):
                    if read[q + i] == ref_seq[r + i]:
                        match += 1
                    else:
                        mism += 1
            elif op in "ID":
                mism += 1
            q += n if op in "MIS=X" else 0
            r += n if op in "MDN=X" else 0
        want = 1.0 if match + mism == 0 else match / (match + mism)
        assert RF.get_gap_compressed_identity_from_alignment(ref_pos0, cigar, bytes(read), ref_seq, True) == want
