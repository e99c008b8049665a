"""Host-side mirror of the reference's rust-bio wrapper (src/bio_align_utils.rs) on top of the CUDA K4 kernel.

Same names and argument order as the Rust: `align(text, pattern)` returns the score, `get_last_alignment()` the
SimpleAlignment built by `bio_alignment_to_simple_alignment` (:16-64), `rescore_alignment` (:71-94), the four
aligner constructors (:97-165) and `OverlapPairwiseAligner` (:207-270).  x = pattern, y = text as in rust-bio.
The reference seeds a band from k-mer matches (k, w); the CUDA path fills the whole matrix: w is accepted and ignored,
k only feeds `find_kmer_matches`, which tells when the two are the same thing (`band_free`).  `align_batch` is the batched form the GPU wants (one launch for many (text, pattern) pairs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import BioScoring
from .batch import PAIR_DTYPE
from .simple_alignment import Cigar, SimpleAlignment

MIN_SCORE = -858993459  # rust-bio's MIN_SCORE: clipping disabled
OP_NAMES = ("Match", "Subst", "Ins", "Del", "Xclip", "Yclip")


@dataclass
class AlignmentWeights:  # src/bio_align_utils.rs:7-12
    match_: int
    mismatch: int
    gap_open: int
    gap_extend: int


@dataclass
class BioAlignment:
    """The fields of bio::alignment::Alignment the reference reads."""
    score: int
    xstart: int
    xend: int
    ystart: int
    yend: int
    operations: List[Tuple[str, int]]  # run-length: (operation, count); clips carry the clipped length


def scoring_from_scores(gap_open: int, gap_extend: int, match: int, mismatch: int) -> BioScoring:
    return BioScoring(gap_open, gap_extend, match, mismatch, MIN_SCORE, MIN_SCORE, MIN_SCORE, MIN_SCORE)


def bio_alignment_to_simple_alignment(aln: BioAlignment) -> SimpleAlignment:
    """:16-64. Match -> '=', Subst -> 'X', Xclip(len) -> soft clip, Yclip dropped; ref_offset = ystart."""
    cigar: List[Cigar] = []
    for op, n in aln.operations:
        if op == "Yclip":
            continue
        if op == "Xclip":
            cigar.append(("S", n))
        else:
            cigar.append(({"Del": "D", "Ins": "I", "Match": "=", "Subst": "X"}[op], n))
    return SimpleAlignment(aln.ystart, cigar)


def rescore_alignment(cigar: Sequence[Cigar], weights: AlignmentWeights) -> int:
    """:71-94."""
    score = 0
    for op, n in cigar:
        if op == "=":
            score += weights.match_ * n
        elif op == "X":
            score += weights.mismatch * n
        elif op == "M":
            raise ValueError("Can't accept match cigar element")
        elif op in "DI":
            score += weights.gap_open + weights.gap_extend * n
    return score


def find_kmer_matches(seq1: bytes, seq2: bytes, k: int) -> List[Tuple[int, int]]:
    """rust-bio `alignment::sparse::find_kmer_matches` as the reference calls it (src/bio_align_utils.rs:169): every
    (i, j) with seq1[i:i+k] == seq2[j:j+k], sorted.  Used here only to tell whether the reference's band degenerates to
    the full matrix (no match at all) — the band construction itself is not restated."""
    if k <= 0 or len(seq1) < k or len(seq2) < k:
        return []
    index = {}
    for i in range(len(seq1) - k + 1):
        index.setdefault(seq1[i:i + k], []).append(i)
    out = []
    for j in range(len(seq2) - k + 1):
        for i in index.get(seq2[j:j + k], ()):
            out.append((i, j))
    out.sort()
    return out


class PairwiseAligner:
    band_free = False  # True after an align() whose pair shares no k-mer: rust-bio fills the whole matrix there as well

    def __init__(self, ctx, scoring: BioScoring, k: int = 20, w: int = 10):
        self.ctx, self.scoring, self.k, self.w = ctx, scoring, k, w
        self.alignment: Optional[BioAlignment] = None

    @classmethod
    def new_merge_aligner(cls, ctx):  # :104-112
        s = scoring_from_scores(0, -2, 1, -3)
        s.yclip_prefix = s.yclip_suffix = 0
        return cls(ctx, s, 20, 10)

    @classmethod
    def new_glocal_aligner(cls, ctx, weights: AlignmentWeights):  # :116-118
        return cls.new_glocal_kw_aligner(ctx, weights, 20, 10)

    @classmethod
    def new_glocal_kw_aligner(cls, ctx, weights: AlignmentWeights, k: int, w: int):  # :122-135
        s = scoring_from_scores(weights.gap_open, weights.gap_extend, weights.match_, weights.mismatch)
        s.yclip_prefix = s.yclip_suffix = 0
        return cls(ctx, s, k, w)

    @classmethod
    def new_pattern_suffix_aligner(cls, ctx, scoring: BioScoring, k: int, w: int):  # :139-146
        s = BioScoring.from_buffer_copy(scoring)
        s.yclip_prefix = 0
        s.xclip_suffix = 0
        return cls(ctx, s, k, w)

    @classmethod
    def new_pattern_prefix_aligner(cls, ctx, scoring: BioScoring, k: int, w: int):  # :150-157
        s = BioScoring.from_buffer_copy(scoring)
        s.yclip_suffix = 0
        s.xclip_prefix = 0
        return cls(ctx, s, k, w)

    def align_batch(self, pairs: Sequence[Tuple[bytes, bytes]]) -> List[BioAlignment]:
        """pairs: (text, pattern) in the reference's argument order."""
        chunks, rows, off = [], [], 0
        for text, pattern in pairs:
            assert len(text) > 0 and len(pattern) > 0  # :177-178
            t = np.frombuffer(bytes(text), np.uint8)
            p = np.frombuffer(bytes(pattern), np.uint8)
            rows.append((off + len(t), off, len(p), len(t), 0, 0, 0, 0, 0))
            chunks += [t, p]
            off += len(t) + len(p)
        arena = np.concatenate(chunks)
        tab = np.array(rows, dtype=PAIR_DTYPE)
        scores, status, bounds, ops, ops_off = self.ctx.bio_align_batch(self.scoring, arena, tab, True)
        assert (status == 0).all(), f"K4 status {status[status != 0][:4]}"
        out = []
        for i in range(len(tab)):
            words = ops[int(ops_off[i]):int(ops_off[i + 1])]
            oplist = [(OP_NAMES[int(w) & 7], int(w) >> 3) for w in words]
            out.append(BioAlignment(int(scores[i]), int(bounds[i, 0]), int(bounds[i, 1]), int(bounds[i, 2]), int(bounds[i, 3]), oplist))
        return out

    def score_batch(self, pairs: Sequence[Tuple[bytes, bytes]]) -> List[int]:
        """Scores only, for callers that never look at the alignment (the merge aligner's: merge_pools.rs:253-257).
        pairs: (text, pattern) in the reference's argument order.  One K5 launch (sfc_bio_score_batch)."""
        chunks, rows, off = [], [], 0
        for text, pattern in pairs:
            assert len(text) > 0 and len(pattern) > 0  # :177-178
            t = np.frombuffer(bytes(text), np.uint8)
            p = np.frombuffer(bytes(pattern), np.uint8)
            rows.append((off + len(t), off, len(p), len(t), 0, 0, 0, 0, 0))
            chunks += [t, p]
            off += len(t) + len(p)
        scores, status = self.ctx.bio_score_batch(self.scoring, np.concatenate(chunks), np.array(rows, dtype=PAIR_DTYPE))
        assert (status == 0).all(), f"K5 status {status[status != 0][:4]}"
        return [int(v) for v in scores]

    def align(self, text: bytes, pattern: bytes) -> int:  # :168-172
        return self.align_with_matches(text, pattern, find_kmer_matches(pattern, text, self.k))

    def align_with_matches(self, text: bytes, pattern: bytes, matches=None) -> int:  # :175-186 (matches only shape the band)
        """`matches` do not change what the GPU computes (full matrix); they decide what can be CLAIMED about it:
        with no k-mer match rust-bio's band is the full matrix too (`band_free` True: identical by construction),
        otherwise the reference result is the optimum inside a band this repo does not restate."""
        self.band_free = matches is not None and len(matches) == 0
        self.alignment = self.align_batch([(text, pattern)])[0]
        return self.alignment.score

    def get_last_alignment(self) -> Optional[SimpleAlignment]:  # :188-191
        return None if self.alignment is None else bio_alignment_to_simple_alignment(self.alignment)


class OverlapPairwiseAligner:
    """Pattern suffix extends the text, or pattern prefix precedes it: two aligners, the better score wins (:207-270)."""

    def __init__(self, ctx, weights: AlignmentWeights):
        s = scoring_from_scores(weights.gap_open, weights.gap_extend, weights.match_, weights.mismatch)
        self.k = 20
        self.is_suffix = True
        self.pattern_suffix_aligner = PairwiseAligner.new_pattern_suffix_aligner(ctx, s, 20, 10)
        self.pattern_prefix_aligner = PairwiseAligner.new_pattern_prefix_aligner(ctx, s, 20, 10)

    @classmethod
    def new_aligner(cls, ctx, weights: AlignmentWeights):
        return cls(ctx, weights)

    def align(self, text: bytes, pattern: bytes) -> int:
        score1 = self.pattern_suffix_aligner.align(text, pattern)
        score2 = self.pattern_prefix_aligner.align(text, pattern)
        self.is_suffix = score1 >= score2
        return max(score1, score2)

    def get_last_alignment(self) -> Optional[SimpleAlignment]:
        return (self.pattern_suffix_aligner if self.is_suffix else self.pattern_prefix_aligner).get_last_alignment()


def test_for_duplicate_haplotypes_with_alignment(ctx, base_query_pairs: Sequence[Tuple[bytes, bytes]],
                                                 min_merge_norm_score: float = 0.97) -> List[bool]:
    """Batched form of the alignment test joint-call uses to merge duplicate haplotypes across samples
    (src/joint_call/merge_haplotypes/single_region/merge_pools.rs:214-272): the query's high-quality range is aligned
    globally against the base contig (text local) with the merge aligner, the score is normalised by the query length,
    and the pair is a duplicate when that is >= 0.97.  base_query_pairs: (base_contig, query_contig[hq_range])."""
    aligner = PairwiseAligner.new_merge_aligner(ctx)
    scores = aligner.score_batch(base_query_pairs)
    return [s / len(q) >= min_merge_norm_score for s, (_b, q) in zip(scores, base_query_pairs)]


test_for_duplicate_haplotypes_with_alignment.__test__ = False  # a reference function name, not a pytest case
