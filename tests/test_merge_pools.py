"""Duplicate-haplotype pooling of joint-call (merge_pools.rs:87-441): host mirror + the two-phase (batched) form.

not gpu: predicates against hand-computed values, two-phase == sequential with the CPU oracle as the aligner.
gpu: the same pools with K5 (sfc_bio_score_batch) as the aligner."""
import numpy as np
import pytest

from oracle import oracle as O
from sawfish_b200 import merge_pools as MP

MERGE = O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0)  # new_merge_aligner, src/bio_align_utils.rs:104-112


def oracle_score(base: bytes, query: bytes) -> int:
    return O.bio_custom_align(MERGE, query, base)[0]  # x = pattern = query, y = text = base


def bp(pos1, pos2, ins=0, dele=None, chrom=0):
    return MP.Breakpoint(MP.Breakend(chrom, pos1), MP.Breakend(chrom, pos2), ins, dele)


def test_breakpoint_predicates():
    assert MP.are_breakends_close(MP.Breakend(0, 100), MP.Breakend(0, 110), 10)
    assert not MP.are_breakends_close(MP.Breakend(0, 100), MP.Breakend(0, 111), 10)
    assert not MP.are_breakends_close(MP.Breakend(0, 100), MP.Breakend(1, 100), 10)
    # insertion sizes: (ins + min(del or 100, 500)); 2|a-b|/(a+b) <= limit
    assert MP.are_breakpoint_insertions_close(bp(0, 0, 0, 0), bp(0, 0, 0, 0), 0.01)            # 0 + 0: trivially close
    assert MP.are_breakpoint_insertions_close(bp(0, 0, 1000, 0), bp(0, 0, 1009, 0), 0.01)      # 18/2009
    assert not MP.are_breakpoint_insertions_close(bp(0, 0, 1000, 0), bp(0, 0, 1011, 0), 0.01)  # 22/2011
    assert MP.are_breakpoint_insertions_close(bp(0, 0, 0, 900), bp(0, 0, 0, 5000), 0.01)       # deletions cap at 500
    assert MP.are_breakpoint_insertions_close(bp(0, 0, 50, None), bp(0, 0, 50, 100), 0.01)     # non-indel counts as 100
    a, b = [bp(1000, 1000, 300, 0)], [bp(1008, 1009, 301, 0)]
    assert MP.are_breakpoint_sets_very_close(a, b)
    assert not MP.are_breakpoint_sets_very_close(a, [bp(1011, 1000, 300, 0)])
    assert not MP.are_breakpoint_sets_very_close(a, b + b)


def test_base_is_the_longer_high_quality_range():
    h1 = MP.MergeHaplotype(b"A" * 50, (5, 45), 3)
    h2 = MP.MergeHaplotype(b"C" * 60, (0, 41), 3)
    assert MP.merge_alignment_pair(h1, h2) == (b"C" * 60, b"A" * 40)   # 41 > 40: hap2 is the base
    h2.high_quality_contig_range = (0, 40)
    assert MP.merge_alignment_pair(h1, h2) == (b"A" * 50, b"C" * 40)   # tie: hap1 stays the base


def synth_pool(rng, n_samples, contig_len=420, n_alleles=3):
    """Samples carrying 1-2 haplotypes each, drawn from a few true alleles of one locus (insertions of different
    sequence/length in the same flanks) with sequencing noise, jittered breakpoints and trimmed high-quality ranges."""
    acgt = np.frombuffer(b"ACGT", np.uint8)
    left, right = acgt[rng.integers(0, 4, contig_len // 2)], acgt[rng.integers(0, 4, contig_len // 2)]
    alleles = []
    for a in range(n_alleles):
        ins_len = int(rng.integers(40, 160))
        alleles.append((np.concatenate([left, acgt[rng.integers(0, 4, ins_len)], right]), ins_len))
    haps, sample_lists = [], []
    for _s in range(n_samples):
        picks = rng.choice(n_alleles, size=int(rng.integers(1, 3)), replace=False)
        lst = []
        for a in picks:
            seq, ins_len = alleles[int(a)]
            seq = seq.copy()
            for _e in range(int(rng.integers(0, 3))):
                seq[int(rng.integers(0, len(seq)))] = acgt[int(rng.integers(0, 4))]
            if rng.random() < 0.3:
                seq = np.delete(seq, int(rng.integers(20, len(seq) - 20)))
            lo = int(rng.integers(0, 30))
            hi = len(seq) - int(rng.integers(0, 30))
            jitter = int(rng.choice([0, 0, 3, 25, 60]))
            pos = 5000 + jitter
            lst.append(len(haps))
            haps.append(MP.MergeHaplotype(seq.tobytes(), (lo, hi), int(rng.integers(2, 30)), [bp(pos, pos, ins_len + int(rng.integers(0, 2)), 0)]))
        sample_lists.append(lst)
    if rng.random() < 0.5:  # a haplotype without SVs is never merged (:381-383)
        haps.append(MP.MergeHaplotype(alleles[0][0].tobytes(), (0, len(alleles[0][0])), 5, []))
        sample_lists.append([len(haps) - 1])
    return haps, sample_lists


def check_pools(pools, haps, sample_lists):
    flat = sorted(h for p in pools for h in p)
    assert flat == list(range(len(haps)))                      # a partition
    for p in pools:
        for lst in sample_lists:                               # two haplotypes of one sample never share a pool
            assert len(set(p) & set(lst)) <= 1
        if len(p) > 1:
            keys = [(-haps[h].supporting_read_count, -len(haps[h].contig_seq)) for h in p]
            assert keys == sorted(keys)


def test_two_phase_equals_sequential_with_the_oracle_aligner():
    rng = np.random.default_rng(41)
    merged_any = False
    for _ in range(6):
        haps, sample_lists = synth_pool(rng, n_samples=int(rng.integers(3, 7)))
        want = MP.get_haplotype_merge_pools_sequential(haps, sample_lists, oracle_score)
        got, counters = MP.get_haplotype_merge_pools(haps, sample_lists, lambda pairs: [oracle_score(b, q) for b, q in pairs])
        assert got == want
        assert counters["alignments_consumed"] <= counters["alignments_batched"]
        check_pools(got, haps, sample_lists)
        merged_any = merged_any or any(len(p) > 1 for p in got)
    assert merged_any


@pytest.mark.gpu
def test_gpu_pools_equal_oracle_pools():
    from sawfish_b200.batch import AlignContext
    ctx = AlignContext(0)
    rng = np.random.default_rng(43)
    score_fn = MP.gpu_merge_score_fn(ctx)
    total = 0
    for _ in range(4):
        haps, sample_lists = synth_pool(rng, n_samples=int(rng.integers(4, 9)), contig_len=int(rng.integers(400, 1500)))
        want = MP.get_haplotype_merge_pools_sequential(haps, sample_lists, oracle_score)
        got, counters = MP.get_haplotype_merge_pools(haps, sample_lists, score_fn)
        assert got == want
        check_pools(got, haps, sample_lists)
        total += counters["alignments_batched"]
    assert total > 0
    ctx.close()


def test_multi_region_pools_two_phase_equals_sequential():
    """multi_region/merge_pools.rs:150-215: greedy pooling of one-haplotype sv groups, breakpoint equality or alignment."""
    rng = np.random.default_rng(47)
    for _ in range(4):
        haps, _lists = synth_pool(rng, n_samples=int(rng.integers(3, 7)))
        groups = [h for h in haps if h.breakpoints]
        got, counters = MP.get_multi_region_merge_pools(groups, lambda pairs: [oracle_score(b, q) for b, q in pairs])
        # the reference's control flow, one alignment call inside the loop
        n = len(groups)
        want = [[x] for x in range(n)]
        for g1 in range(n):
            if not want[g1]:
                continue
            for g2 in range(g1 + 1, n):
                if not want[g2]:
                    continue
                if groups[g1].breakpoints[:1] == groups[g2].breakpoints[:1]:
                    is_dup = True
                else:
                    base, query = MP.merge_alignment_pair(groups[g1], groups[g2])
                    is_dup = oracle_score(base, query) / len(query) >= 0.97
                if is_dup:
                    want[g1].extend(want[g2])
                    want[g2] = []
        for p in want:
            if len(p) > 1:
                p.sort(key=lambda g: (-groups[g].supporting_read_count, -len(groups[g].contig_seq)))
        assert got == want
        assert sorted(g for p in got for g in p) == list(range(n))
        assert counters["alignments_consumed"] <= counters["alignments_batched"]
