"""Joint-call's ingest of one sample's `discover` output directory, as far as the alignment path needs it:
`contig.alignment.bam` (assembled contigs + their alignments, get_candidate_sv_contig_info) and `candidate.sv.bcf` (candidate
SV records -> refined SVs, get_anno_refined_svs) — reference src/joint_call/get_refined_svs.rs:176-195,554-601.  The result is
what the merge aligner's caller (merge_pools.py) and score_sv's allele-flank extraction (read_flanks.py) start from."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List

from . import candidate_sv as CS
from . import contig_bam as CB
from . import merge_pools as MP


@dataclass
class SampleDiscoverData:
    sample_assemblies: List[List[CB.ClusterAssembly]]     # [cluster_index][assembly_index]
    refined_svs: List[CS.AnnotatedRefinedSV]


def load_discover_dir(discover_dir: str, label_to_index: Dict[str, int]) -> SampleDiscoverData:
    contig_path = os.path.join(discover_dir, CB.CONTIG_ALIGNMENT_FILENAME)
    sv_path = os.path.join(discover_dir, CS.CANDIDATE_SV_FILENAME)
    if not os.path.exists(contig_path):
        raise FileNotFoundError(f"Can't find contig alignment file expected at {contig_path}")   # get_refined_svs.rs:183-185
    if not os.path.exists(sv_path):
        raise FileNotFoundError(f"Can't find candidate SV file expected at {sv_path}")            # :566-568
    assemblies = CB.get_candidate_sv_contig_info(contig_path)
    svs = []
    for rec in CS.read_candidate_bcf(sv_path, label_to_index):
        a = CS.process_record_to_anno_refined_sv(label_to_index, rec)
        if a is not None:
            svs.append(a)
    return SampleDiscoverData(assemblies, svs)


def merge_haplotypes_of_samples(samples: List[SampleDiscoverData]):
    """One MergeHaplotype per (sample, cluster, assembly) that carries at least one refined SV, with the breakpoints of its
    SVs (in record order) — the view get_haplotype_merge_pools works on.  Returns (haplotypes, sample_haplotype_list)."""
    haps: List[MP.MergeHaplotype] = []
    sample_lists: List[List[int]] = []
    for s in samples:
        by_hap: Dict[tuple, List[MP.Breakpoint]] = {}
        for sv in s.refined_svs:
            bp = sv.bp
            dele = bp.deletion_size()
            mbp = MP.Breakpoint(MP.Breakend(bp.breakend1.chrom_index, bp.breakend1.start), MP.Breakend(bp.breakend2.chrom_index, bp.breakend2.start),
                                len(bp.insert_seq), dele)
            by_hap.setdefault((sv.id.cluster_index, sv.id.assembly_index), []).append(mbp)
        lst = []
        for (ci, ai), bps in sorted(by_hap.items()):
            asm = s.sample_assemblies[ci][ai]
            a0 = asm.contig_alignments[0]
            lst.append(len(haps))
            haps.append(MP.MergeHaplotype(a0.contig_seq, a0.high_quality_contig_range, asm.supporting_read_count, bps))
        sample_lists.append(lst)
    return haps, sample_lists


# ---------------------------------------------------------------------------------------------- refined SVs -> SV groups
@dataclass
class SVGroup:
    """The fields of crate::sv_group::SVGroup that get_refined_svs.rs fills for a single sample."""
    cluster_index: int
    group_haplotypes: List[CB.ClusterAssembly]      # kept haplotypes, re-indexed
    sample_haplotype_list: List[List[int]]          # [sample] -> haplotype indexes (one sample here)
    sv_haplotype_map: List[int]                     # refined SV -> haplotype index
    refined_svs: List[CS.AnnotatedRefinedSV]


def process_rsv_cluster_to_sv_group(contigs: List[List[CB.ClusterAssembly]], group_arsvs: List[CS.AnnotatedRefinedSV],
                                    max_haplotype_count: int) -> "SVGroup | None":
    """process_rsv_cluster_to_sv_group (get_refined_svs.rs:606-768) for one cluster's refined SVs.  The sample's haplotype
    list is the first SV's contig map cut to `max_haplotype_count` (get_max_haplotype_count_for_regions,
    src/expected_ploidy.rs:146-168: 2 for a single-region diploid locus, 1 for haploid loci and for multi-region SVs);
    haplotypes outside that list are dropped together with the SVs that sit on them, and every index is re-mapped."""
    assert group_arsvs
    first = group_arsvs[0]
    cluster_index = first.id.cluster_index
    group_haplotypes = list(contigs[cluster_index])
    sample_haplotype_list = [list(first.contig_map[:max_haplotype_count])]
    refined_svs = list(group_arsvs)
    sv_haplotype_map = [x.id.assembly_index for x in refined_svs]
    if len(sample_haplotype_list[0]) < len(group_haplotypes):
        remap = [None] * len(group_haplotypes)
        for new_index, old_index in enumerate(sample_haplotype_list[0]):
            remap[old_index] = new_index
        keep_sv = [remap[h] is not None for h in sv_haplotype_map]
        sample_haplotype_list[0] = [remap[h] for h in sample_haplotype_list[0]]
        assert all(h is not None for h in sample_haplotype_list[0])
        sv_haplotype_map = [remap[h] for h in sv_haplotype_map if remap[h] is not None]
        group_haplotypes = [x for i, x in enumerate(group_haplotypes) if remap[i] is not None]
        refined_svs = [x for x, k in zip(refined_svs, keep_sv) if k]
    if not refined_svs:
        return None
    return SVGroup(cluster_index, group_haplotypes, sample_haplotype_list, sv_haplotype_map, refined_svs)


def group_single_sample_refined_svs(contigs: List[List[CB.ClusterAssembly]], anno_refined_svs: List[CS.AnnotatedRefinedSV],
                                    max_haplotype_count_of_cluster=lambda cluster_index: 2) -> List[SVGroup]:
    """group_single_sample_refined_svs (:796-846): sort by (cluster, assembly, alignment) index, one SV group per cluster."""
    svs = sorted(anno_refined_svs, key=lambda x: (x.id.cluster_index, x.id.assembly_index, x.id.alignment_index))
    groups: List[SVGroup] = []
    cur: List[CS.AnnotatedRefinedSV] = []
    for sv in svs + [None]:
        if cur and (sv is None or sv.id.cluster_index != cur[0].id.cluster_index):
            g = process_rsv_cluster_to_sv_group(contigs, cur, max_haplotype_count_of_cluster(cur[0].id.cluster_index))
            if g is not None:
                groups.append(g)
            cur = []
        if sv is not None:
            cur.append(sv)
    return groups
