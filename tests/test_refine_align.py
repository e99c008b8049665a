"""refine_sv side: CIGAR walker on the reference's CIGAR shapes (CPU) and planted-SV recovery through the GPU."""
import numpy as np
import pytest

from sawfish_b200.refine_sv_align import get_single_region_indel_candidates
from sawfish_b200.simple_alignment import SimpleAlignment, cigar_from_string


def test_cigar_walker_anchor_rules():
    al = SimpleAlignment(10, cigar_from_string("100=40D200=60I80=5X30=36D20="))
    got = get_single_region_indel_candidates(al, chrom_pos=1010, high_quality_range=(0, 10 ** 6))
    # 40D and 60I are anchored by >=50 '=' runs on both sides; the trailing 36D has no closing anchor (20= < 50)
    assert [(c.del_len, c.ins_len, c.ref_pos, c.contig_pos) for c in got] == [(40, 0, 1110, 100), (0, 60, 1350, 300)]
    # events outside the high-quality contig range are ignored
    assert get_single_region_indel_candidates(al, 1010, (0, 50)) == []
    # below min_indel_size
    al2 = SimpleAlignment(0, cigar_from_string("100=34D100="))
    assert get_single_region_indel_candidates(al2, 0, (0, 1000)) == []
    # first event needs an opening anchor too
    al3 = SimpleAlignment(0, cigar_from_string("20=40D100=50I100="))
    assert [(c.del_len, c.ins_len) for c in get_single_region_indel_candidates(al3, 0, (0, 1000))] == [(0, 50)]
    # large-insertion mode keeps only insertions > 3000
    al4 = SimpleAlignment(0, cigar_from_string("100=3500I100=40D100="))
    assert [(c.ins_len) for c in get_single_region_indel_candidates(al4, 0, (0, 10 ** 6), is_large_insertion=True)] == [3500]


@pytest.mark.gpu
def test_planted_sv_recovered_through_gpu_alignment(ctx):
    from sawfish_b200 import synth
    from sawfish_b200.refine_sv_align import align_contigs_batch
    rng = np.random.default_rng(8)
    targets, contigs, truth = [], [], []
    for i in range(12):
        window = synth.random_seq(rng, 7600)
        hap = window[1500:6100]
        L = int(rng.integers(40, 1200))
        pos = int(rng.integers(1700, 2700))
        if i % 2:
            contig = np.concatenate([hap[:pos], synth.random_seq(rng, L), hap[pos:]])
            truth.append(("I", L, 1500 + pos))
        else:
            contig = np.concatenate([hap[:pos], hap[pos + L:]])
            truth.append(("D", L, 1500 + pos))
        targets.append(window.tobytes())
        contigs.append((i, contig.tobytes()))
    res = align_contigs_batch(ctx, targets, contigs)
    for (score, al), (kind, L, ref_pos), (ti, c) in zip(res, truth, contigs):
        assert al is not None
        cands = get_single_region_indel_candidates(al, al.ref_offset, (0, len(c)))
        assert len(cands) == 1
        cand = cands[0]
        assert (cand.ins_len, cand.del_len) == ((L, 0) if kind == "I" else (0, L))
        assert abs(cand.ref_pos - ref_pos) <= 12   # placement inside the indel's homology range is re-normalised downstream
