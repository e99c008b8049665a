"""Synthetic HiFi-like workloads for the alignment hot path (BASELINE.json configs 2, 2b, 3, 5).

All generators are seeded and return (arena uint8 ASCII, pairs PAIR_DTYPE) in exactly the form the C ABI
takes.  Pairs carry the sawfish convention: both sequences reversed, clip = min(0.4|p|, 0.4|t|) on all four
ends (reference src/wfa2_utils.rs:106-118,198-215).
"""
from __future__ import annotations

import numpy as np

from .batch import PAIR_DTYPE, PAIR_REVERSE, sawfish_clip

_ACGT = np.frombuffer(b"ACGT", np.uint8)


def random_seq(rng: np.random.Generator, n: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, n)]


def add_errors(rng: np.random.Generator, seq: np.ndarray, rate: float) -> np.ndarray:
    """i.i.d. errors: 1/3 substitution, 1/3 1-bp insertion, 1/3 1-bp deletion."""
    n = len(seq)
    ne = rng.binomial(n, rate) if rate > 0 else 0
    if ne == 0:
        return seq.copy()
    pos = np.sort(rng.choice(n, ne, replace=False))
    kinds = rng.integers(0, 3, ne)
    out = []
    prev = 0
    for p, k in zip(pos, kinds):
        out.append(seq[prev:p])
        if k == 0:  # substitution (to a different base)
            b = _ACGT[(np.searchsorted(_ACGT, seq[p]) + rng.integers(1, 4)) % 4]
            out.append(np.array([b], np.uint8))
        elif k == 1:  # insertion before p
            out.append(np.array([_ACGT[rng.integers(0, 4)], seq[p]], np.uint8))
        # k == 2: deletion of seq[p]
        prev = p + 1
    out.append(seq[prev:])
    return np.concatenate(out)


def plant_indel(rng: np.random.Generator, hap: np.ndarray, length: int, margin: int) -> np.ndarray:
    """hap with one insertion or deletion (50/50) of `length` at an interior position."""
    n = len(hap)
    margin = min(margin, max(1, n // 4))
    if rng.random() < 0.5 or n - 2 * margin - length <= 0:
        pos = int(rng.integers(margin, max(margin + 1, n - margin)))
        return np.concatenate([hap[:pos], random_seq(rng, length), hap[pos:]])
    pos = int(rng.integers(margin, n - margin - length))
    return np.concatenate([hap[:pos], hap[pos + length:]])


class _Builder:
    def __init__(self):
        self.chunks = []
        self.off = 0
        self.rows = []

    def add_seq(self, s: np.ndarray) -> int:
        o = self.off
        self.chunks.append(np.ascontiguousarray(s, np.uint8))
        self.off += len(s)
        return o

    def add_pair(self, p_off, p_len, t_off, t_len, clip=None, reverse=True):
        c = sawfish_clip(p_len, t_len) if clip is None else clip
        self.rows.append((p_off, t_off, p_len, t_len, c, c, c, c, PAIR_REVERSE if reverse else 0))

    def finish(self):
        arena = np.concatenate(self.chunks) if self.chunks else np.zeros(0, np.uint8)
        pairs = np.array(self.rows, dtype=PAIR_DTYPE)
        return arena, pairs


def log_uniform_int(rng, lo, hi):
    return int(round(np.exp(rng.uniform(np.log(lo), np.log(hi)))))


def config2_read_vs_haplotypes(n_reads: int, seed: int = 0x5AF15 + 2, len_lo: int = 2000, len_hi: int = 20000,
                               sv_lo: int = 35, sv_hi: int = 5000, err: float = 0.001):
    """BASELINE config 2: each read segment (text) vs 2 allele haplotypes (patterns).

    hap0 random, hap1 = hap0 + one indel SV; read = noisy copy of hap0 or hap1. 2*n_reads pairs.
    With len 100-500 / sv 35-300 this is config 2b, the shape score_sv really issues
    (reference src/score_sv/mod.rs:837-901: text = read flank, pattern = allele flank).
    """
    rng = np.random.default_rng(seed)
    b = _Builder()
    truth = []
    for _ in range(n_reads):
        L = int(rng.integers(len_lo, len_hi + 1))
        hap0 = random_seq(rng, L)
        sv = min(log_uniform_int(rng, sv_lo, sv_hi), max(1, L // 2))
        hap1 = plant_indel(rng, hap0, sv, 500 if L >= 2000 else L // 5)
        which = int(rng.integers(0, 2))
        read = add_errors(rng, hap1 if which else hap0, err)
        t_off = b.add_seq(read)
        for hap in (hap0, hap1):
            p_off = b.add_seq(hap)
            b.add_pair(p_off, len(hap), t_off, len(read))
        truth.append(which)
    arena, pairs = b.finish()
    return arena, pairs, np.array(truth, np.int8)


def config2b_score_sv_flanks(n_reads: int, seed: int = 0x5AF15 + 12):
    return config2_read_vs_haplotypes(n_reads, seed, 100, 500, 35, 300, 0.001)


def config3_refine_contigs(n_clusters: int, seed: int = 0x5AF15 + 3, two_region_frac: float = 0.2):
    """BASELINE config 3: assembled contig (pattern) vs reference window (text), refine_sv-shaped.

    Single-region: contig = 1500 flank + ~600 core + SV + 1500 flank vs window = target +- 3500
    (reference src/refine_sv/mod.rs:663-664, src/refine_sv/align/mod.rs:394-404).
    Two-region: text = seg1 + 100 x 'N' + seg2 (src/refine_sv/align/mod.rs:712-737), contig spans the junction.
    """
    rng = np.random.default_rng(seed)
    b = _Builder()
    for _ in range(n_clusters):
        if rng.random() < two_region_frac:
            seg1 = random_seq(rng, int(rng.integers(3500, 5500)))
            seg2 = random_seq(rng, int(rng.integers(3500, 5500)))
            text = np.concatenate([seg1, np.full(100, ord("N"), np.uint8), seg2])
            flank = int(rng.choice([1500, 2500]))
            contig = np.concatenate([seg1[-flank - 300:], seg2[:flank + 300]])
            contig = add_errors(rng, contig, 0.0005)
        else:
            target = int(rng.integers(200, 1500))
            window = random_seq(rng, target + 7000)
            core_lo = 3500 - 1500 - 300
            core_hi = 3500 + target + 1500 + 300
            hap = window[core_lo:core_hi]
            sv = log_uniform_int(rng, 35, 5000)
            contig = plant_indel(rng, hap, min(sv, len(hap) // 3), 1600)
            contig = add_errors(rng, contig, 0.0005)
            text = window
        t_off = b.add_seq(text)
        p_off = b.add_seq(contig)
        b.add_pair(p_off, len(contig), t_off, len(text))
    arena, pairs = b.finish()
    return arena, pairs


def config5_stress(n_pairs: int, seed: int = 0x5AF15 + 5, len_lo: int = 10000, len_hi: int = 50000,
                   div_lo: float = 0.01, div_hi: float = 0.05):
    """BASELINE config 5: large insertion/duplication haplotypes at high divergence."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    for _ in range(n_pairs):
        L = int(rng.integers(len_lo, len_hi + 1))
        t = random_seq(rng, L)
        ins = int(rng.integers(len_lo, len_hi + 1))
        pos = int(rng.integers(L // 4, 3 * L // 4))
        if rng.random() < 0.5:  # tandem duplication of the preceding segment (as far as it exists)
            dup = t[max(0, pos - ins):pos]
            p = np.concatenate([t[:pos], dup, t[pos:]])
        else:
            p = np.concatenate([t[:pos], random_seq(rng, ins), t[pos:]])
        p = add_errors(rng, p, float(rng.uniform(div_lo, div_hi)))
        t_off = b.add_seq(t)
        p_off = b.add_seq(p)
        b.add_pair(p_off, len(p), t_off, len(t))
    return b.finish()


def random_pairs(n: int, seed: int, max_len: int = 120, alphabet: bytes = b"ACGTN", reverse_frac: float = 0.5,
                 sawfish_clip_frac: float = 0.5):
    """Adversarial small pairs for parity tests: low-complexity alphabets, random ends-free allowances,
    related and unrelated sequences (tie-breaks bite on homopolymers / tandem repeats)."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    rows = []
    for _ in range(n):
        alpha = np.frombuffer(alphabet, np.uint8)
        k = int(rng.integers(1, len(alpha) + 1))
        alpha = alpha[rng.permutation(len(alpha))[:k]]
        tl = int(rng.integers(1, max_len + 1))
        t = alpha[rng.integers(0, len(alpha), tl)]
        mode = rng.random()
        if mode < 0.7:
            p = t.copy()
            for _ in range(int(rng.integers(0, 4))):
                if len(p) == 0:
                    break
                a = int(rng.integers(0, len(p) + 1))
                r = rng.random()
                if r < 0.4:
                    p = np.concatenate([p[:a], alpha[rng.integers(0, len(alpha), int(rng.integers(1, 40)))], p[a:]])
                elif r < 0.8:
                    e = int(rng.integers(a, min(len(p), a + 40) + 1))
                    p = np.concatenate([p[:a], p[e:]])
                else:
                    p = add_errors(rng, p, 0.1)
            if len(p) == 0:
                p = alpha[:1].copy()
        else:
            p = alpha[rng.integers(0, len(alpha), int(rng.integers(1, max_len + 1)))]
        pl, tl = len(p), len(t)
        t_off = b.add_seq(t)
        p_off = b.add_seq(p)
        if rng.random() < sawfish_clip_frac:
            c = sawfish_clip(pl, tl)
            ends = (c, c, c, c)
        elif rng.random() < 0.2:
            ends = (0, 0, 0, 0)
        else:
            ends = (int(rng.integers(0, pl + 1)), int(rng.integers(0, pl + 1)), int(rng.integers(0, tl + 1)),
                    int(rng.integers(0, tl + 1)))
        rows.append((p_off, t_off, pl, tl) + ends + (PAIR_REVERSE if rng.random() < reverse_frac else 0,))
    arena = np.concatenate(b.chunks)
    return arena, np.array(rows, dtype=PAIR_DTYPE)


def gcups_cells(pairs: np.ndarray) -> int:
    """Full-matrix DP cells sum |p|*|t| (the GCUPS-equivalent numerator of BASELINE.json's metric)."""
    return int((pairs["pattern_len"].astype(np.int64) * pairs["text_len"].astype(np.int64)).sum())


def low_complexity_pairs(n: int, seed: int, len_lo: int = 300, len_hi: int = 3000):
    """Adversarial mid-size pairs where backtrace tie-breaks bite: tandem repeats, homopolymer runs and 2-letter
    sequences with repeat-unit indels (many co-optimal alignments), N runs, plus random ends-free allowances."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    rows = []
    for _ in range(n):
        L = int(rng.integers(len_lo, len_hi + 1))
        kind = int(rng.integers(0, 4))
        if kind == 0:    # tandem repeat of a short unit with point noise
            unit = random_seq(rng, int(rng.integers(1, 7)))
            t = np.tile(unit, L // len(unit) + 1)[:L].copy()
            noise = rng.random(L) < 0.01
            t[noise] = _ACGT[rng.integers(0, 4, int(noise.sum()))]
        elif kind == 1:  # 2-letter random
            t = np.frombuffer(b"AC", np.uint8)[rng.integers(0, 2, L)]
        elif kind == 2:  # homopolymer blocks
            blocks = []
            while sum(len(x) for x in blocks) < L:
                blocks.append(np.full(int(rng.integers(1, 40)), _ACGT[rng.integers(0, 4)], np.uint8))
            t = np.concatenate(blocks)[:L]
        else:            # random with an N run (the refine_sv spacer)
            t = random_seq(rng, L)
            a = int(rng.integers(0, max(1, L - 100)))
            t[a:a + 100] = ord("N")
        p = t.copy()
        for _ in range(int(rng.integers(1, 4))):   # repeat-scale indels
            a = int(rng.integers(0, len(p)))
            n_ev = int(rng.integers(1, 60))
            if rng.random() < 0.5:
                p = np.concatenate([p[:a], p[max(0, a - n_ev):a], p[a:]])   # local duplication
            else:
                p = np.concatenate([p[:a], p[a + n_ev:]])
        if len(p) == 0:
            p = t[:1].copy()
        p = add_errors(rng, p, 0.003)
        t_off = b.add_seq(t)
        p_off = b.add_seq(p)
        pl, tl = len(p), len(t)
        if rng.random() < 0.6:
            c = sawfish_clip(pl, tl)
            ends = (c, c, c, c)
        else:
            ends = (int(rng.integers(0, pl // 2 + 1)), int(rng.integers(0, pl // 2 + 1)), int(rng.integers(0, tl // 2 + 1)),
                    int(rng.integers(0, tl // 2 + 1)))
        rows.append((p_off, t_off, pl, tl) + ends + (PAIR_REVERSE if rng.random() < 0.5 else 0,))
    arena = np.concatenate(b.chunks)
    return arena, np.array(rows, dtype=PAIR_DTYPE)
