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
