"""Two-region alt-haplotype text (reference src/refine_sv/align/mod.rs:708-1330) and the speculative multi-region batch
(:1975-2199 in two-phase form)."""
import numpy as np
import pytest

from sawfish_b200 import two_region_alt_hap as T
from sawfish_b200.contig_bam import rev_comp


def _ref(n=60000, seed=0):
    rng = np.random.default_rng(seed)
    return {0: bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]), 1: bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 9000)])}, [n, 9000]


def test_construct_alt_hap_seq_orders_and_orientations():
    s1, s2 = b"AAAACCCC", b"GGTT"
    assert T.construct_alt_hap_seq(s1, s2, 3, False, True) == (b"AAAACCCCNNNGGTT", (8, 11))
    assert T.construct_alt_hap_seq(s1, s2, 3, False, False) == (b"GGTTNNNAAAACCCC", (4, 7))
    assert T.construct_alt_hap_seq(s1, s2, 3, True, True) == (b"AAAACCCCNNN" + rev_comp(s2), (8, 11))
    assert T.construct_alt_hap_seq(s1, s2, 0, True, False) == (rev_comp(s2) + s1, (4, 4))


def test_flank_size_rules():
    # interior flank (next to the spacer) is capped at 2000, the outer one keeps the full size (:776-793)
    assert T.get_left_right_ref_flank_size(3500, False, True) == (3500, 2000)
    assert T.get_left_right_ref_flank_size(3500, False, False) == (2000, 3500)
    assert T.get_left_right_ref_flank_size(3500, True, True) == (2000, 3500)
    assert T.get_left_right_ref_flank_size(1500, True, False) == (1500, 1500)
    segs = [T.GenomeSegment(0, 1000, 1100), T.GenomeSegment(0, 2100, 2200)]
    dele = T.Breakpoint(T.Breakend(segs[0], T.LEFT_ANCHOR), T.Breakend(segs[1], T.RIGHT_ANCHOR))
    inv = T.Breakpoint(T.Breakend(segs[0], T.LEFT_ANCHOR), T.Breakend(segs[1], T.LEFT_ANCHOR))
    assert T.get_middle_flank_size(3500, segs, dele) == 500          # half of the 1000 bases between the regions (:755-770)
    assert T.get_middle_flank_size(3500, segs, inv) == 3500           # same orientation: no restriction
    other = [segs[0], T.GenomeSegment(1, 2100, 2200)]
    assert T.get_middle_flank_size(3500, other, dele) == 3500         # different chromosomes: no restriction


def test_deletion_like_breakpoint_layout_and_pins():
    ref, sizes = _ref()
    segs = [T.GenomeSegment(0, 9990, 10010), T.GenomeSegment(0, 15990, 16010)]
    bp = T.Breakpoint(T.Breakend(T.GenomeSegment(0, 10000, 10001), T.LEFT_ANCHOR), T.Breakend(T.GenomeSegment(0, 16000, 16001), T.RIGHT_ANCHOR))
    info = T.get_two_region_alt_hap_info(ref, sizes, segs, bp)
    assert (info.ref_segment1.start, info.ref_segment1.end) == (9990 - 3500, 10010 + 2000)
    assert (info.ref_segment2.start, info.ref_segment2.end) == (15990 - 2000, 16010 + 3500)
    a, b = info.spacer_range
    assert info.alt_hap_seq[:a] == ref[0][info.ref_segment1.start:info.ref_segment1.end] and info.alt_hap_seq[a:b] == b"N" * 100
    assert info.alt_hap_seq[b:] == ref[0][info.ref_segment2.start:info.ref_segment2.end]
    assert not info.ref_segment2_revcomp and info.ref_segment2_after_ref_segment1
    assert not info.left_boundary_pin and not info.right_boundary_pin
    assert info.segment1_neighbor_extension_size == 0 and len(info.extended_alt_hap_seq) == 2 * 7020 + 100
    # close regions: the flank between them shrinks to half the distance so that the segments do not overlap
    segs2 = [T.GenomeSegment(0, 9990, 10010), T.GenomeSegment(0, 10810, 10830)]
    info2 = T.get_two_region_alt_hap_info(ref, sizes, segs2, bp)
    assert info2.ref_segment1.end == 10010 + 400 and info2.ref_segment2.start == 10810 - 400
    # a segment at the chromosome start: the clipped flank pins the boundary
    segs3 = [T.GenomeSegment(1, 0, 20), T.GenomeSegment(0, 15990, 16010)]
    bp3 = T.Breakpoint(T.Breakend(T.GenomeSegment(1, 10, 11), T.RIGHT_ANCHOR), T.Breakend(T.GenomeSegment(0, 16000, 16001), T.LEFT_ANCHOR))
    info3 = T.get_two_region_alt_hap_info(ref, sizes, segs3, bp3)
    assert not info3.ref_segment2_after_ref_segment1 and info3.right_boundary_pin and not info3.left_boundary_pin


def test_inversion_like_breakpoint_revcomps_segment2():
    ref, sizes = _ref()
    segs = [T.GenomeSegment(0, 9990, 10010), T.GenomeSegment(0, 30990, 31010)]
    bp = T.Breakpoint(T.Breakend(T.GenomeSegment(0, 10000, 10001), T.LEFT_ANCHOR), T.Breakend(T.GenomeSegment(0, 31000, 31001), T.LEFT_ANCHOR))
    info = T.get_two_region_alt_hap_info(ref, sizes, segs, bp)
    assert info.ref_segment2_revcomp and info.ref_segment2_after_ref_segment1
    b = info.spacer_range[1]
    assert info.alt_hap_seq[b:] == rev_comp(ref[0][info.ref_segment2.start:info.ref_segment2.end])
    # revcomped and second: its outer flank is on the left in reference space (:776-793)
    assert (info.ref_segment2.start, info.ref_segment2.end) == (30990 - 3500, 31010 + 2000)


def test_neighbor_breakend_extension():
    """A neighbouring breakpoint 1200 bases left of a left-anchored target breakend: the left flank stops there and the
    remaining 2300 bases of alt-haplotype flank come from the neighbour's remote side (:822-956)."""
    ref, sizes = _ref()
    seg = T.GenomeSegment(0, 20000, 20020)
    target = T.Breakend(T.GenomeSegment(0, 20010, 20011), T.LEFT_ANCHOR)
    local = T.Breakend(T.GenomeSegment(0, 18800, 18801), T.RIGHT_ANCHOR)
    remote = T.Breakend(T.GenomeSegment(0, 40000, 40001), T.LEFT_ANCHOR)
    nb = T.FullBreakendNeighborInfo(0, T.Breakpoint(local, remote))
    lf, rf, rseq = T.get_remote_breakend_info(3500, ref, sizes, seg, target, nb)
    assert (lf, rf) == (1200, 3500) and len(rseq) == 2300 and rseq == ref[0][40000 - 2300:40000]
    s, seq, ext, al, ar = T.extract_reference_segment_info(3500, ref, sizes, seg, target, 3500, 2000, nb)
    assert (s.start, s.end) == (18800, 22020) and len(seq) == 2300 + 1200 + 20 + 2000 and seq.startswith(rseq)
    assert (al, ar) == (1200, 3500) and len(ext) == 2300 + 1200 + 20 + 3500


@pytest.mark.gpu
def test_speculative_multi_region_batch_on_gpu(ctx, oracle):
    """Contigs of several breakpoint clusters (1500- and 2500-flank variants, some too noisy to pass) in ONE batch; the
    selection afterwards equals aligning contig by contig in the reference's order, and the CIGAR of the accepted contig
    crosses the N spacer with a pure deletion (segment_alignment.rs:26-58)."""
    from sawfish_b200 import segment_alignment as SA
    from sawfish_b200 import synth
    from sawfish_b200.wfa2_utils import PairwiseAligner
    ref, sizes = _ref(seed=5)
    rng = np.random.default_rng(9)
    batch = T.MultiRegionBatch()
    clusters = []
    for ci in range(6):
        p1 = 8000 + 7000 * ci
        p2 = p1 + int(rng.integers(3000, 6000))
        segs = [T.GenomeSegment(0, p1 - 10, p1 + 10), T.GenomeSegment(0, p2 - 10, p2 + 10)]
        bp = T.Breakpoint(T.Breakend(T.GenomeSegment(0, p1, p1 + 1), T.LEFT_ANCHOR), T.Breakend(T.GenomeSegment(0, p2, p2 + 1), T.RIGHT_ANCHOR))
        info = T.get_two_region_alt_hap_info(ref, sizes, segs, bp)
        asms = []
        for ai in range(2):
            contigs = []
            for flank in (1500, 2500):
                c = np.frombuffer(ref[0][p1 - flank - 300:p1] + ref[0][p2:p2 + flank + 300], np.uint8)
                noise = 0.2 if (flank == 1500 and (ci + ai) % 3 == 0) else 0.0005     # some first contigs are junk
                contigs.append(synth.add_errors(rng, c, noise).tobytes())
            asms.append(T.AssemblyContigs(contigs, 10))
        batch.add_cluster(ci, info, asms)
        clusters.append((info, asms))
    batch.align(ctx)
    al = PairwiseAligner.new_large_indel_aligner(False, ctx)
    picked_second = 0
    for ci, (info, asms) in enumerate(clusters):
        al.set_text(info.alt_hap_seq)
        for ai, asm in enumerate(asms):
            def accept(_k, a):
                return SA.does_alignment_support_alt_hap(50, a, info.spacer_range[0], info.spacer_range[1])
            got = batch.first_passing(ci, ai, len(asm.contig_seqs), accept)
            want = None
            for k, seq in enumerate(asm.contig_seqs):       # the reference's sequential loop
                score, a = al.align(seq)
                assert (score, a) == batch.alignment(ci, ai, k)
                if a is not None and accept(k, a):
                    want = (k, a)
                    break
            assert got == want and got is not None
            picked_second += got[0]
    assert picked_second > 0          # the retry contig was needed somewhere
