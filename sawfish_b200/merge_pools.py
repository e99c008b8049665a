"""Host-side mirror of joint-call's duplicate-haplotype pooling, with the alignment test batched for the GPU.

Reference: src/joint_call/merge_haplotypes/single_region/merge_pools.rs
  are_breakends_close (:87-90), are_breakpoint_insertions_close (:97-120), are_breakpoints_close (:124-144),
  are_breakpoint_sets_very_close (:146-160), test_for_duplicate_haplotypes_with_alignment (:214-272),
  test_for_duplicate_haplotypes (:276-323), get_haplotype_merge_pools (:333-441).

The reference walks the haplotype pairs (h1 < h2) of a merged SV group sequentially and calls the merge aligner (rust-bio,
`PairwiseAligner::new_merge_aligner`) inside the loop; which pairs are reached depends on earlier merges (a merged
haplotype's pool is emptied, exclusion sets are inherited).  The alignment result itself depends only on the two contigs,
so the loop is split the same way as score_sv's:
  phase 1  enumerate every pair the loop COULD reach and that the breakpoint shortcuts do not settle, build
           (base contig, query high-quality range) for each;
  phase 2  ONE sfc_bio_score_batch (K5, score only) -> norm score = score / len(query) >= 0.97;
  phase 3  replay the reference's sequential loop against the precomputed answers.
Pairs that the sequential loop would have skipped cost one speculative alignment each; nothing about the result changes.
Only the fields of Breakpoint / SVGroup that this path reads are modelled.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

MIN_MERGE_NORM_SCORE = 0.97  # :220


@dataclass(frozen=True)
class Breakend:  # the fields of crate::breakpoint::Breakend this path reads
    chrom_index: int
    start: int


@dataclass(frozen=True)
class Breakpoint:
    breakend1: Breakend
    breakend2: Breakend
    insert_size: int = 0                 # bp.insert_info.size()
    deletion_size: Optional[int] = None  # bp.deletion_size(): None for non-indel breakpoints


@dataclass
class MergeHaplotype:
    contig_seq: bytes                        # contig_info.contig_alignments[0].contig_seq
    high_quality_contig_range: Tuple[int, int]
    supporting_read_count: int
    breakpoints: List[Breakpoint] = field(default_factory=list)  # the bps of the refined SVs mapped to this haplotype


def are_breakends_close(be1: Breakend, be2: Breakend, max_breakend_distance: int) -> bool:  # :87-90
    return be1.chrom_index == be2.chrom_index and abs(be1.start - be2.start) <= max_breakend_distance


def are_breakpoint_insertions_close(bp1: Breakpoint, bp2: Breakpoint, max_insert_size_diff: float) -> bool:  # :97-120
    non_indel_size, max_del_size = 100, 500
    del1 = min(bp1.deletion_size if bp1.deletion_size is not None else non_indel_size, max_del_size)
    del2 = min(bp2.deletion_size if bp2.deletion_size is not None else non_indel_size, max_del_size)
    size1, size2 = bp1.insert_size + del1, bp2.insert_size + del2
    if size1 + size2 == 0:
        return True
    return (2 * abs(size1 - size2)) / (size1 + size2) <= max_insert_size_diff


def are_breakpoints_close(bp1: Breakpoint, bp2: Breakpoint, max_breakend_distance: int, max_insert_size_diff: float) -> bool:  # :124-144
    if not are_breakends_close(bp1.breakend1, bp2.breakend1, max_breakend_distance):
        return False
    if not are_breakends_close(bp1.breakend2, bp2.breakend2, max_breakend_distance):
        return False
    return are_breakpoint_insertions_close(bp1, bp2, max_insert_size_diff)


def are_breakpoint_sets_very_close(bps1: Sequence[Breakpoint], bps2: Sequence[Breakpoint]) -> bool:  # :146-160
    return len(bps1) == len(bps2) and all(are_breakpoints_close(a, b, 10, 0.01) for a, b in zip(bps1, bps2))


def merge_alignment_pair(hap1: MergeHaplotype, hap2: MergeHaplotype) -> Tuple[bytes, bytes]:
    """(base contig, query contig[hq range]) of test_for_duplicate_haplotypes_with_alignment (:222-251): the haplotype
    with the LONGER high-quality range is the base (hap1 on ties), the other one's high-quality range is the query."""
    size1 = hap1.high_quality_contig_range[1] - hap1.high_quality_contig_range[0]
    size2 = hap2.high_quality_contig_range[1] - hap2.high_quality_contig_range[0]
    base, query = (hap2, hap1) if size2 > size1 else (hap1, hap2)
    lo, hi = query.high_quality_contig_range
    return base.contig_seq, query.contig_seq[lo:hi]


def breakpoint_shortcut(bps1: Sequence[Breakpoint], bps2: Sequence[Breakpoint]) -> bool:
    """The two tests of test_for_duplicate_haplotypes that need no alignment (:290-298)."""
    return list(bps1) == list(bps2) or are_breakpoint_sets_very_close(bps1, bps2)


ScoreFn = Callable[[List[Tuple[bytes, bytes]]], List[int]]   # [(base, query)] -> merge-aligner scores


def gpu_merge_score_fn(ctx) -> ScoreFn:
    """One K5 launch per call (sfc_bio_score_batch with the merge scoring)."""
    from .bio_align_utils import PairwiseAligner
    aligner = PairwiseAligner.new_merge_aligner(ctx)
    return lambda pairs: aligner.score_batch(pairs) if pairs else []


def get_haplotype_merge_pools(haplotypes: Sequence[MergeHaplotype], sample_haplotype_list: Sequence[Sequence[int]],
                              score_fn: ScoreFn) -> Tuple[List[List[int]], Dict[str, int]]:
    """:333-441 in two-phase form.  Returns (haplotype_merge_pools, counters)."""
    n = len(haplotypes)
    # exclusion sets: haplotypes of one sample never merge (:350-366)
    excluded = [set() for _ in range(n)]
    for lst in sample_haplotype_list:
        if len(lst) > 1:
            for h in lst:
                excluded[h].update(x for x in lst if x != h)
    # ---- phase 1: every pair the sequential loop could reach and the shortcuts do not settle ----
    # Exclusion sets only grow and pools only empty during the loop, so the statically excluded / empty-breakpoint pairs
    # are never reached; everything else is speculated.
    need: Dict[Tuple[int, int], int] = {}
    batch: List[Tuple[bytes, bytes]] = []
    for h1 in range(n):
        if not haplotypes[h1].breakpoints:
            continue
        for h2 in range(h1 + 1, n):
            if not haplotypes[h2].breakpoints or h2 in excluded[h1]:
                continue
            if breakpoint_shortcut(haplotypes[h1].breakpoints, haplotypes[h2].breakpoints):
                continue
            need[(h1, h2)] = len(batch)
            batch.append(merge_alignment_pair(haplotypes[h1], haplotypes[h2]))
    # ---- phase 2: one batch ----
    scores = score_fn(batch)
    dup = {k: (scores[i] / len(batch[i][1])) >= MIN_MERGE_NORM_SCORE for k, i in need.items()}
    # ---- phase 3: the reference's loop, verbatim, against the precomputed answers ----
    pools: List[List[int]] = [[x] for x in range(n)]
    used = 0
    for h1 in range(n):
        if not pools[h1]:
            continue
        if not haplotypes[h1].breakpoints:
            continue
        for h2 in range(h1 + 1, n):
            if not pools[h2]:
                continue
            if not haplotypes[h2].breakpoints:
                continue
            if h2 in excluded[h1]:
                continue
            if breakpoint_shortcut(haplotypes[h1].breakpoints, haplotypes[h2].breakpoints):
                is_dup = True
            else:
                is_dup = dup[(h1, h2)]
                used += 1
            if is_dup:
                pools[h1].extend(pools[h2])
                pools[h2] = []
                if excluded[h2]:
                    excluded[h1].update(excluded[h2])
                    excluded[h2] = set()
    # representative first: highest read support, then longest contig (:425-439; sort_by_key is stable)
    for pool in pools:
        if len(pool) > 1:
            pool.sort(key=lambda h: (-haplotypes[h].supporting_read_count, -len(haplotypes[h].contig_seq)))
    return pools, {"alignments_batched": len(batch), "alignments_consumed": used}


def get_haplotype_merge_pools_sequential(haplotypes: Sequence[MergeHaplotype], sample_haplotype_list: Sequence[Sequence[int]],
                                         score_one: Callable[[bytes, bytes], int]) -> List[List[int]]:
    """The reference's control flow as written (one alignment call inside the loop); used by the tests to show that the
    two-phase form returns the same pools."""
    n = len(haplotypes)
    pools = [[x] for x in range(n)]
    excluded = [set() for _ in range(n)]
    for lst in sample_haplotype_list:
        if len(lst) > 1:
            for h in lst:
                excluded[h].update(x for x in lst if x != h)
    for h1 in range(n):
        if not pools[h1] or not haplotypes[h1].breakpoints:
            continue
        for h2 in range(h1 + 1, n):
            if not pools[h2] or not haplotypes[h2].breakpoints or h2 in excluded[h1]:
                continue
            if breakpoint_shortcut(haplotypes[h1].breakpoints, haplotypes[h2].breakpoints):
                is_dup = True
            else:
                base, query = merge_alignment_pair(haplotypes[h1], haplotypes[h2])
                is_dup = score_one(base, query) / len(query) >= MIN_MERGE_NORM_SCORE
            if is_dup:
                pools[h1].extend(pools[h2])
                pools[h2] = []
                if excluded[h2]:
                    excluded[h1].update(excluded[h2])
                    excluded[h2] = set()
    for pool in pools:
        if len(pool) > 1:
            pool.sort(key=lambda h: (-haplotypes[h].supporting_read_count, -len(haplotypes[h].contig_seq)))
    return pools


# ---------------------------------------------------------------------------------------------- multi-region SV groups
def get_multi_region_merge_pools(groups: Sequence[MergeHaplotype], score_fn: ScoreFn) -> Tuple[List[List[int]], Dict[str, int]]:
    """The pooling step of merge_duplicated_candidates for multi-region (breakpoint) SV groups
    (src/joint_call/merge_haplotypes/multi_region/merge_pools.rs:150-215), in the same two-phase form: every sv group holds
    exactly one SV and one haplotype; two groups are duplicates when their breakpoints are EQUAL (:86-88) or the merge aligner
    says so (:12-75, same base / query rule and 0.97 threshold as the single-region case); no exclusion sets here."""
    n = len(groups)
    need: Dict[Tuple[int, int], int] = {}
    batch: List[Tuple[bytes, bytes]] = []
    for g1 in range(n):
        for g2 in range(g1 + 1, n):
            if list(groups[g1].breakpoints[:1]) == list(groups[g2].breakpoints[:1]):
                continue
            need[(g1, g2)] = len(batch)
            batch.append(merge_alignment_pair(groups[g1], groups[g2]))
    scores = score_fn(batch)
    dup = {k: (scores[i] / len(batch[i][1])) >= MIN_MERGE_NORM_SCORE for k, i in need.items()}
    pools: List[List[int]] = [[x] for x in range(n)]
    used = 0
    for g1 in range(n):
        if not pools[g1]:
            continue
        for g2 in range(g1 + 1, n):
            if not pools[g2]:
                continue
            if (g1, g2) in dup:
                is_dup = dup[(g1, g2)]
                used += 1
            else:
                is_dup = True
            if is_dup:
                pools[g1].extend(pools[g2])
                pools[g2] = []
    for pool in pools:
        if len(pool) > 1:
            pool.sort(key=lambda g: (-groups[g].supporting_read_count, -len(groups[g].contig_seq)))
    return pools, {"alignments_batched": len(batch), "alignments_consumed": used}
