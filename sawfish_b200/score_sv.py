"""Batched read-support scoring: the caller side of the hot path in joint-call.

Mirrors, with the same names and tie-break behaviour, the pieces of the reference's src/score_sv/mod.rs that sit
directly on either side of the alignment call:

  get_read_flank_align_score_option   :721-745   (allele-flank truncation + min-flank filter -> pair or None)
  get_read_flank_align_score          :695-713   (norm_score = score / text.len())
  get_read_support_scores_from_segment :837-901  (per read x breakend x registration: ref, alt, overlap alleles)
       -> here in TWO PHASES: enumerate every slot of the SV group into one pair batch, align it in one call
          (consensus preset, score only), scatter the scores back
  sort_best_allele_score_conditions   :951-960
  get_best_allele_for_flank_condition :970-1063
  check_if_read_allele_depth_count_is_usable :1073-1115
  set_allele_depth_counts_for_single_sample_and_sv_allele :1130-1369 (the per-read loop body)
  get_sv_sample_qualities_from_allele_counts :1422-1521, get_sv_qualities_from_allele_counts :1527-1549
  get_haploid_ln_prior / get_diploid_ln_prior :1972-1995; src/prob_utils.rs:7-21

NOT mirrored (out of the path's scope): BAM fetch and read filters, flank extraction geometry from BAM records and
contig alignments, SVGroup bookkeeping (guest haplotypes, phasing).  The synthetic SV groups below supply flank
sequences directly; everything downstream of the integer scores is the reference's logic.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from enum import IntEnum
from functools import cmp_to_key
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .batch import CONSENSUS_PENALTIES, PAIR_DTYPE, PAIR_REVERSE, sawfish_clip


class AlleleType(IntEnum):  # src/refine_sv/mod.rs:80-84
    Ref = 0
    Alt = 1
    Overlap = 2


class Genotype(IntEnum):  # src/refine_sv/mod.rs:105-109
    Ref = 0
    Het = 1
    Hom = 2


BREAKEND1, BREAKEND2 = 0, 1            # BreakendFlankType
STANDARD, ASM_REGION_EDGE = 0, 1       # BreakendFlankRegistrationType

# ScoreSVSettings::new (src/score_sv/mod.rs:2036-2060) and cli default max_qscore (src/cli/joint_call.rs:63-64)
SV_PRIOR = 5e-4
READ_SUPPORT_FLANK_SIZE = 500
MIN_READ_SUPPORT_FLANK_SIZE = 100
MAX_QSCORE = 999
F64_MIN_10_EXP = -307.0


# ---------------------------------------------------------------------------------------- prob utils
def _rust_round(x: float) -> int:
    """f64::round (half away from zero) followed by `as i32` (saturating)."""
    if math.isnan(x):
        return 0
    r = math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)
    return int(max(-2 ** 31, min(2 ** 31 - 1, r)))


def error_prob_to_phred(prob: float) -> float:  # src/prob_utils.rs:7-9
    lg = math.log10(prob) if prob > 0 else -math.inf
    return -10.0 * max(lg, F64_MIN_10_EXP)


def ln_error_prob_to_qphred(ln_prob: float) -> int:  # src/prob_utils.rs:11-21
    return _rust_round(-10.0 * max(ln_prob / math.log(10.0), F64_MIN_10_EXP))


def normalize_ln_distro(x: List[float]) -> int:  # src/prob_utils.rs:27-52
    max_index, max_p = 0, x[0]
    for i, p in enumerate(x[1:]):
        if p > max_p:
            max_p, max_index = p, i + 1
    e = [math.exp(p - max_p) if p != -math.inf else 0.0 for p in x]
    s = sum(e)
    for i in range(len(x)):
        x[i] = e[i] / s
    return max_index


def get_haploid_ln_prior(prior: float) -> List[float]:  # :1972-1982
    p = [0.0, 0.0, 0.0]
    p[Genotype.Hom] = prior
    p[Genotype.Ref] = 1.0 - p[Genotype.Hom]
    return [math.log(v) if v > 0 else -math.inf for v in p]


def get_diploid_ln_prior(prior: float) -> List[float]:  # :1984-1995
    p = [0.0, 0.0, 0.0]
    p[Genotype.Het] = prior
    p[Genotype.Hom] = prior / 2.0
    p[Genotype.Ref] = 1.0 - (p[Genotype.Het] + p[Genotype.Hom])
    return [math.log(v) for v in p]


# ---------------------------------------------------------------------------------------- score containers
@dataclass
class AlleleScores:
    """breakend_flanks[flank_type][registration] = Option<f64> (src/score_sv/mod.rs:1371-1402)."""
    breakend_flanks: List[List[Optional[float]]] = field(default_factory=lambda: [[None, None], [None, None]])


@dataclass
class OverlappingHaplotypeSupportScores:
    haplotype_index: int
    scores: AlleleScores = field(default_factory=AlleleScores)


@dataclass
class ReadSupportScores:
    ref_allele: AlleleScores = field(default_factory=AlleleScores)
    alt_allele: AlleleScores = field(default_factory=AlleleScores)
    overlapping_alleles: List[OverlappingHaplotypeSupportScores] = field(default_factory=list)


@dataclass
class FlankAlleleScore:
    score: Optional[float]
    allele_type: AlleleType


@dataclass
class BestAlleleInfo:
    best_allele_score: FlankAlleleScore
    tying_score_allele_type: Optional[AlleleType]
    flank_type: int = BREAKEND1
    flank_registration_type: int = STANDARD


def _option_cmp(a: Optional[float], b: Optional[float]) -> int:
    """Option<f64>::partial_cmp: None < Some(_)."""
    if a is None or b is None:
        return (a is not None) - (b is not None)
    return (a > b) - (a < b)


def sort_best_allele_score_conditions(a: FlankAlleleScore, b: FlankAlleleScore) -> int:  # :951-960
    if a.allele_type != AlleleType.Ref and b.allele_type == AlleleType.Ref:
        return -1
    if b.allele_type != AlleleType.Ref and a.allele_type == AlleleType.Ref:
        return 1
    return _option_cmp(b.score, a.score)


def sort_best_allele_conditions(a: BestAlleleInfo, b: BestAlleleInfo) -> int:  # :962-965
    return sort_best_allele_score_conditions(a.best_allele_score, b.best_allele_score)


def get_best_allele_for_flank_condition(read_scores: ReadSupportScores, overlapping_allele_index: Optional[int],
                                        is_alt_on_top_haplotype_rank: bool, flank_type: int, flank_registration_type: int,
                                        read_to_hap_allele_assignment: Optional[AlleleType]) -> BestAlleleInfo:  # :970-1063
    min_flank_score = 0.0
    allele_scores = [
        FlankAlleleScore(read_scores.ref_allele.breakend_flanks[flank_type][flank_registration_type], AlleleType.Ref),
        FlankAlleleScore(read_scores.alt_allele.breakend_flanks[flank_type][flank_registration_type], AlleleType.Alt),
    ]
    if overlapping_allele_index is not None:
        allele_scores.append(FlankAlleleScore(
            read_scores.overlapping_alleles[overlapping_allele_index].scores.breakend_flanks[flank_type][flank_registration_type],
            AlleleType.Overlap))
    allele_scores.sort(key=cmp_to_key(lambda a, b: _option_cmp(b.score, a.score)))  # stable, like slice::sort_by
    tying = allele_scores[1].allele_type if allele_scores[0].score == allele_scores[1].score else None
    best_allele_index = 0
    if tying is not None and allele_scores[0].allele_type != AlleleType.Ref and allele_scores[1].allele_type != AlleleType.Ref:
        if read_to_hap_allele_assignment is not None:
            if allele_scores[1].allele_type == read_to_hap_allele_assignment:
                best_allele_index = 1
        elif (not is_alt_on_top_haplotype_rank and allele_scores[0].allele_type == AlleleType.Alt
              and allele_scores[1].allele_type == AlleleType.Overlap):
            best_allele_index = 1
    if best_allele_index == 1:
        tying = allele_scores[0].allele_type
    x = FlankAlleleScore(allele_scores[best_allele_index].score, allele_scores[best_allele_index].allele_type)
    if x.score is not None and x.score < min_flank_score:
        x.score = None
    return BestAlleleInfo(x, tying, flank_type, flank_registration_type)


def check_if_read_allele_depth_count_is_usable(is_assigned_allele_score_tie: bool, on_multiple_sample_haplotypes: bool,
                                               assigned_allele: AlleleType, qname: str, sv_haplotype_index: int,
                                               overlapping_haplotype_index: Optional[int],
                                               read_to_hap: Dict[str, int]) -> bool:  # :1073-1115
    if is_assigned_allele_score_tie or on_multiple_sample_haplotypes:
        return True
    if assigned_allele == AlleleType.Alt:
        return read_to_hap.setdefault(qname, sv_haplotype_index) == sv_haplotype_index
    if assigned_allele == AlleleType.Overlap:
        if overlapping_haplotype_index is not None:
            return read_to_hap.setdefault(qname, overlapping_haplotype_index) == overlapping_haplotype_index
        return False
    return True


def set_allele_depth_counts(locus_scores: Dict[str, ReadSupportScores], sv_haplotype_index: int,
                            overlapping_haplotype_index: Optional[int], is_alt_on_top_haplotype_rank: bool,
                            on_multiple_sample_haplotypes: bool, read_to_hap: Dict[str, int]) -> List[int]:
    """Per-read loop of set_allele_depth_counts_for_single_sample_and_sv_allele (:1177-1355).
    Returns allele_depth indexed by AlleleType.  locus_scores is iterated in qname order (BTreeMap, :1598)."""
    allele_depth = [0, 0, 0]
    for qname in sorted(locus_scores):
        read_scores = locus_scores[qname]
        overlapping_allele_index = None
        if overlapping_haplotype_index is not None:
            for i, x in enumerate(read_scores.overlapping_alleles):
                if x.haplotype_index == overlapping_haplotype_index:
                    overlapping_allele_index = i
                    break
        read_to_hap_allele_assignment = None
        if overlapping_allele_index is not None and qname in read_to_hap:
            read_to_hap_allele_assignment = AlleleType.Alt if read_to_hap[qname] == sv_haplotype_index else AlleleType.Overlap
        tests = [get_best_allele_for_flank_condition(read_scores, overlapping_allele_index, is_alt_on_top_haplotype_rank, ft, rt,
                                                     read_to_hap_allele_assignment)
                 for ft in (BREAKEND1, BREAKEND2) for rt in (STANDARD, ASM_REGION_EDGE)]
        tests.sort(key=cmp_to_key(sort_best_allele_conditions))
        best = 0
        tying_allele_type = tests[0].tying_score_allele_type
        if tying_allele_type is not None:
            for idx in range(1, len(tests)):
                if tests[idx].tying_score_allele_type is None:
                    if tests[idx].best_allele_score.score is not None and tests[idx].best_allele_score.allele_type == tying_allele_type:
                        best = idx
                    break
        assigned = tests[best]
        if assigned.best_allele_score.score is None:
            continue
        assigned_allele = assigned.best_allele_score.allele_type
        is_tie = assigned.tying_score_allele_type is not None
        count_allele = assigned_allele
        if assigned_allele == AlleleType.Overlap:
            count_allele = AlleleType.Alt if on_multiple_sample_haplotypes else AlleleType.Overlap
        if check_if_read_allele_depth_count_is_usable(is_tie, on_multiple_sample_haplotypes, assigned_allele, qname,
                                                      sv_haplotype_index, overlapping_haplotype_index, read_to_hap):
            allele_depth[count_allele] += 1
    return allele_depth


@dataclass
class SampleQuality:
    gt: Optional[Genotype]
    gq: int
    pl: List[int]
    ad: List[int]          # VCF AD = [Ref + Overlap, Alt]  (src/large_variant_output.rs:443-463)
    ref_pprob: float


def get_sv_sample_qualities_from_allele_counts(allele_depth: List[int], is_haploid: bool = False,
                                               max_qscore: int = MAX_QSCORE) -> SampleQuality:  # :1422-1521
    allele_genotype_probs = [[1.0, 0.0], [0.5, 0.5], [0.0, 1.0]]
    wrong_read_allele_lhood = 1e-5
    gt_ln_lhoods = [0.0, 0.0, 0.0]
    for g in range(3):
        if is_haploid and g == Genotype.Het:
            gt_ln_lhoods[g] = -math.inf
            continue
        ap = allele_genotype_probs[g]
        for read_allele_index, count in enumerate(allele_depth):
            lh = [1.0, 1.0]
            lh[0 if read_allele_index == 1 else 1] = wrong_read_allele_lhood  # Alt reads: Ref is the wrong allele
            read_lhood = lh[0] * ap[0] + lh[1] * ap[1]
            gt_ln_lhoods[g] += math.log(read_lhood) * float(count)
    prior = get_haploid_ln_prior(SV_PRIOR) if is_haploid else get_diploid_ln_prior(SV_PRIOR)
    gt_pprob = [gt_ln_lhoods[g] + prior[g] for g in range(3)]
    # ".enumerate().rev().max_by(total_cmp)": max_by returns the LAST maximum of the reversed iterator = lowest index
    max_gt_lhood_index = 0
    for g in (1, 2):
        if gt_ln_lhoods[g] > gt_ln_lhoods[max_gt_lhood_index]:
            max_gt_lhood_index = g
    normalize_ln_distro(gt_pprob)
    pl = [min(ln_error_prob_to_qphred(gt_ln_lhoods[g] - gt_ln_lhoods[max_gt_lhood_index]), max_qscore) for g in range(3)]
    gt = Genotype(max_gt_lhood_index) if sum(allele_depth) != 0 else None
    gq = min(pl[g] for g in range(3) if g != max_gt_lhood_index)
    ad = [allele_depth[AlleleType.Ref] + allele_depth[AlleleType.Overlap], allele_depth[AlleleType.Alt]]
    return SampleQuality(gt, gq, pl, ad, gt_pprob[Genotype.Ref])


def get_sv_qualities_from_allele_counts(sample_allele_depths: List[List[int]], max_qscore: int = MAX_QSCORE):  # :1527-1549
    joint_ref_prob = 1.0
    samples = []
    for ad in sample_allele_depths:
        q = get_sv_sample_qualities_from_allele_counts(ad, False, max_qscore)
        samples.append(q)
        joint_ref_prob *= q.ref_pprob
    qual = max(0.0, min(float(np.float32(error_prob_to_phred(joint_ref_prob))), float(max_qscore)))
    return samples, qual


# ---------------------------------------------------------------------------------------- two-phase batching
@dataclass
class ReadFlank:
    """One read x breakend x registration: the read flank (text) and how much of the flank window was truncated by
    the read ends (breakend_flank.start_truncation / end_truncation, src/score_sv/mod.rs:863-864)."""
    seq: np.ndarray
    start_truncation: int = 0
    end_truncation: int = 0


@dataclass
class ScoreSlot:
    qname: str
    breakend: int
    registration: int
    allele: AlleleType
    overlap_index: int      # index into overlapping haplotype list, -1 otherwise
    text_len: int


class PairBatchBuilder:
    """Phase 1: enumerate every (read flank, allele flank) slot of an SV group into one arena + pair table."""

    def __init__(self):
        self.chunks: List[np.ndarray] = []
        self.off = 0
        self.rows: List[tuple] = []
        self.slots: List[Tuple[int, ScoreSlot]] = []   # (group tag, slot)

    def _add(self, s: np.ndarray) -> int:
        o = self.off
        self.chunks.append(np.ascontiguousarray(s, np.uint8))
        self.off += len(s)
        return o

    def add_read_flank(self, tag: int, qname: str, breakend: int, registration: int, flank: ReadFlank,
                       ref_seq: Optional[np.ndarray], alt_seq: Optional[np.ndarray],
                       overlap_seqs: List[Optional[np.ndarray]]):
        """== the body of the registration loop at src/score_sv/mod.rs:829-901 minus the alignment itself."""
        t_off = None
        for allele, oi, seq in [(AlleleType.Ref, -1, ref_seq), (AlleleType.Alt, -1, alt_seq)] + \
                               [(AlleleType.Overlap, i, s) for i, s in enumerate(overlap_seqs)]:
            # get_read_flank_align_score_option (:721-745)
            if seq is None:
                continue
            if flank.start_truncation + flank.end_truncation + MIN_READ_SUPPORT_FLANK_SIZE > len(seq):
                continue
            seq = seq[flank.start_truncation:len(seq) - flank.end_truncation]
            if t_off is None:
                t_off = self._add(flank.seq)
            p_off = self._add(seq)
            c = sawfish_clip(len(seq), len(flank.seq))
            self.rows.append((p_off, t_off, len(seq), len(flank.seq), c, c, c, c, PAIR_REVERSE))
            self.slots.append((tag, ScoreSlot(qname, breakend, registration, allele, oi, len(flank.seq))))

    def finish(self):
        arena = np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.uint8)
        return arena, np.array(self.rows, dtype=PAIR_DTYPE)


def scatter_scores(slots: List[Tuple[int, ScoreSlot]], scores: np.ndarray, overlap_haplotype_indexes: Dict[int, List[int]]):
    """Phase 3: norm_score = score / text.len() (:702-703) written into ReadSupportScores keyed by (tag, qname)."""
    out: Dict[int, Dict[str, ReadSupportScores]] = {}
    for (tag, sl), sc in zip(slots, scores):
        rs = out.setdefault(tag, {}).setdefault(sl.qname, ReadSupportScores())
        haps = overlap_haplotype_indexes.get(tag, [])
        while len(rs.overlapping_alleles) < len(haps):
            rs.overlapping_alleles.append(OverlappingHaplotypeSupportScores(haps[len(rs.overlapping_alleles)]))
        norm = float(int(sc)) / float(sl.text_len)
        if sl.allele == AlleleType.Ref:
            rs.ref_allele.breakend_flanks[sl.breakend][sl.registration] = norm
        elif sl.allele == AlleleType.Alt:
            rs.alt_allele.breakend_flanks[sl.breakend][sl.registration] = norm
        else:
            rs.overlapping_alleles[sl.overlap_index].scores.breakend_flanks[sl.breakend][sl.registration] = norm
    return out


# ---------------------------------------------------------------------------------------- synthetic joint-call
def synth_sv_groups(n_groups: int, n_samples: int, depth: int, seed: int):
    """BASELINE config 4 in miniature: per SV group one deletion/insertion allele (plus, for some groups, a second
    overlapping haplotype); per sample a genotype and `depth` reads, each giving a <=500 bp flank at both breakends in
    one or two registrations.  Returns (builder, meta) ready for one alignment batch."""
    from . import synth
    rng = np.random.default_rng(seed)
    b = PairBatchBuilder()
    overlap_haps: Dict[int, List[int]] = {}
    truth = {}
    half = READ_SUPPORT_FLANK_SIZE // 2
    for gi in range(n_groups):
        ref = synth.random_seq(rng, 4000)
        sv_len = synth.log_uniform_int(rng, 50, 800)
        pos = 2000
        if rng.random() < 0.5:   # deletion allele
            alt = np.concatenate([ref[:pos], ref[pos + sv_len:]])
            alt_bps, ref_bps = (pos, pos), (pos, pos + sv_len)
        else:                    # insertion allele
            alt = np.concatenate([ref[:pos], synth.random_seq(rng, sv_len), ref[pos:]])
            alt_bps, ref_bps = (pos, pos + sv_len), (pos, pos)
        has_overlap = rng.random() < 0.3
        ovl = None
        if has_overlap:          # a second haplotype with a different nearby deletion
            l2 = synth.log_uniform_int(rng, 50, 400)
            ovl = np.concatenate([ref[:pos - 40], ref[pos - 40 + l2:]])
            ovl_bps = (pos - 40, pos - 40)
        for si in range(n_samples):
            tag = gi * n_samples + si
            gt = int(rng.choice([0, 1, 2], p=[0.5, 0.35, 0.15]))
            truth[tag] = gt
            overlap_haps[tag] = [1] if has_overlap else []
            for ri in range(depth):
                from_alt = rng.random() < (0.0, 0.5, 1.0)[gt]
                from_ovl = has_overlap and (not from_alt) and rng.random() < 0.3
                hap, bps = (alt, alt_bps) if from_alt else ((ovl, ovl_bps) if from_ovl else (ref, ref_bps))
                qname = f"read{ri:04d}"
                for be in (BREAKEND1, BREAKEND2):
                    center = bps[be]
                    lo, hi = center - half, center + half
                    # the read covers the flank window, sometimes only partly (truncation by the read end)
                    st = int(rng.integers(0, 120)) if rng.random() < 0.15 else 0
                    et = int(rng.integers(0, 120)) if rng.random() < 0.15 else 0
                    read_flank = synth.add_errors(rng, hap[lo + st:hi - et], 0.001)
                    def cut(h, c):
                        return h[c - half:c + half]
                    regs = [STANDARD] + ([ASM_REGION_EDGE] if rng.random() < 0.2 else [])
                    for reg in regs:
                        b.add_read_flank(tag, qname, be, reg, ReadFlank(read_flank, st, et), cut(ref, ref_bps[be]),
                                         cut(alt, alt_bps[be]), [cut(ovl, ovl_bps[be])] if has_overlap else [])
    return b, {"overlap_haps": overlap_haps, "truth": truth, "n_groups": n_groups, "n_samples": n_samples}


AlignFn = Callable[[np.ndarray, np.ndarray], np.ndarray]   # (arena, pairs) -> int32 scores


def joint_call_score_groups(builder: PairBatchBuilder, meta: dict, align_fn: AlignFn):
    """Phase 2 (one alignment batch) + phase 3 (scatter, AD, GT/GQ/PL, QUAL) for every (group, sample)."""
    arena, pairs = builder.finish()
    scores = align_fn(arena, pairs)
    locus = scatter_scores(builder.slots, scores, meta["overlap_haps"])
    results = {}
    for gi in range(meta["n_groups"]):
        ads = []
        for si in range(meta["n_samples"]):
            tag = gi * meta["n_samples"] + si
            haps = meta["overlap_haps"][tag]
            ad = set_allele_depth_counts(locus.get(tag, {}), sv_haplotype_index=0,
                                         overlapping_haplotype_index=haps[0] if len(haps) == 1 else None,
                                         is_alt_on_top_haplotype_rank=True, on_multiple_sample_haplotypes=False, read_to_hap={})
            ads.append(ad)
        samples, qual = get_sv_qualities_from_allele_counts(ads)
        results[gi] = (samples, qual)
    return results, len(pairs)


def gpu_align_fn(ctx) -> AlignFn:
    def f(arena, pairs):
        scores, status, _, _ = ctx.align_batch(CONSENSUS_PENALTIES, arena, pairs, False)
        assert (status == 0).all()   # == assert_eq!(status, StatusAlgCompleted), src/wfa2_utils.rs:117
        return scores
    return f
