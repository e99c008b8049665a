"""rust-bio style aligner (bio_align_utils): oracle vs the reference's known answers and an independent definition
(not gpu), CUDA K4 vs oracle and the same known answers through the host mirror (gpu).

Known answers: reference src/bio_align_utils.rs:291-403 (nine of them, four aligner flavours)."""
import numpy as np
import pytest

from oracle import oracle as O

TEXT = b"AACTGTATAA"


def rle(ops):
    out = []
    for name, n in ops:
        if out and out[-1][0] == name and name not in ("Xclip", "Yclip"):
            out[-1] = (name, out[-1][1] + n)
        else:
            out.append((name, n))
    return out


def cigar_of(ops):
    m = {"Del": "D", "Ins": "I", "Match": "=", "Subst": "X", "Xclip": "S"}
    return "".join(f"{n}{m[name]}" for name, n in rle(ops) if name != "Yclip")


def scorings():
    w = dict(gap_open=-1, gap_extend=-1, match=1, mismatch=-3)  # get_test_weights, :281-288
    return {
        "merge": O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0),
        "glocal": O.bio_scoring(**w, yclip_prefix=0, yclip_suffix=0),
        "suffix": O.bio_scoring(**w, yclip_prefix=0, xclip_suffix=0),
        "prefix": O.bio_scoring(**w, yclip_suffix=0, xclip_prefix=0),
    }


KNOWN = [  # flavour, pattern, score, (ref_offset, cigar) or None when the reference only checks the score
    ("merge", b"ACTTAT", 4, (1, "3=1D3=")),
    ("glocal", b"CTGAATA", 3, (2, "3=1X3=")),
    ("glocal", b"CCGAACT", 0, None),
    ("glocal", b"ACTGATA", 5, (1, "4=1D3=")),
    ("suffix", b"ATAACCG", 4, (6, "4=3S")),
    ("suffix", b"CCGAACT", 0, None),
    ("prefix", b"CCGAACT", 4, (0, "3S4=")),
    ("prefix", b"ATAACCG", 1, None),
]


@pytest.mark.parametrize("flavour,pattern,score,aln", KNOWN)
def test_oracle_reproduces_reference_known_answers(flavour, pattern, score, aln):
    s, xs, xe, ys, ye, ops = O.bio_custom_align(scorings()[flavour], pattern, TEXT)
    assert s == score
    if aln is not None:
        assert (ys, cigar_of(ops)) == aln


def global_affine(sc, x, y):
    """Plain Gotoh global alignment score (gap of length L costs open + L * extend), independent of the oracle's code."""
    NEG = -10**9
    m, n = len(x), len(y)
    S = [[NEG] * (n + 1) for _ in range(m + 1)]
    I = [[NEG] * (n + 1) for _ in range(m + 1)]
    D = [[NEG] * (n + 1) for _ in range(m + 1)]
    S[0][0] = 0
    for i in range(m + 1):
        for j in range(n + 1):
            if i > 0:
                I[i][j] = max(I[i - 1][j] + sc.gap_extend, S[i - 1][j] + sc.gap_open + sc.gap_extend)
            if j > 0:
                D[i][j] = max(D[i][j - 1] + sc.gap_extend, S[i][j - 1] + sc.gap_open + sc.gap_extend)
            if i > 0 and j > 0:
                S[i][j] = S[i - 1][j - 1] + (sc.match if x[i - 1] == y[j - 1] else sc.mismatch)
            if i or j:
                S[i][j] = max(S[i][j], I[i][j], D[i][j])
    return S[m][n]


def test_oracle_glocal_score_is_best_text_substring():
    """yclip(0): pattern global, text local == max over text substrings of the global affine score."""
    rng = np.random.default_rng(5)
    sc = scorings()["merge"]
    for _ in range(12):
        x = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 8))))
        y = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 10))))
        want = max(global_affine(sc, x, y[a:b]) for a in range(len(y) + 1) for b in range(a, len(y) + 1))
        got = O.bio_custom_align(sc, x, y)[0]
        assert got == want, (x, y, got, want)


def test_oracle_operations_replay_to_the_score():
    rng = np.random.default_rng(6)
    for name, sc in scorings().items():
        for _ in range(30):
            base = rng.choice(list(b"ACGT"), int(rng.integers(5, 60)))
            x = bytes(base[int(rng.integers(0, 3)):])
            y = bytes(rng.choice(list(b"ACGT"), int(rng.integers(0, 6)))) + bytes(base) + bytes(rng.choice(list(b"ACGT"), 3))
            s, xs, xe, ys, ye, ops = O.bio_custom_align(sc, x, y)
            i, j, tot, prev = 0, 0, 0, None
            for op, n in ops:
                if op in ("Match", "Subst"):
                    assert (x[i] == y[j]) == (op == "Match")
                    tot += sc.match if op == "Match" else sc.mismatch
                    i += 1; j += 1
                elif op == "Ins":
                    tot += sc.gap_extend + (sc.gap_open if prev != "Ins" else 0); i += 1
                elif op == "Del":
                    tot += sc.gap_extend + (sc.gap_open if prev != "Del" else 0); j += 1
                elif op == "Xclip":
                    tot += sc.xclip_prefix if i == 0 else sc.xclip_suffix; i += n
                else:
                    tot += sc.yclip_prefix if j == 0 else sc.yclip_suffix; j += n
                prev = op
            assert (i, j, tot) == (len(x), len(y), s), (name, x, y, ops)


# ---------------------------------------------------------------------------------------------------------- GPU
def _to_gpu_scoring(sc):
    from sawfish_b200._lib import BioScoring
    return BioScoring(sc.gap_open, sc.gap_extend, sc.match, sc.mismatch, sc.xclip_prefix, sc.xclip_suffix, sc.yclip_prefix,
                      sc.yclip_suffix)


def test_find_kmer_matches_is_every_shared_kmer_sorted():
    """`find_kmer_matches` (what `align` hands to rust-bio, src/bio_align_utils.rs:169) against a brute-force scan; with no
    shared k-mer the reference's band is the full matrix, which is where this repo's full-matrix result is identical by
    construction (`band_free`)."""
    import numpy as np
    from sawfish_b200.bio_align_utils import find_kmer_matches
    rng = np.random.default_rng(5)
    for k in (3, 5, 20):
        x = bytes(rng.choice(list(b"ACGT"), 300).tolist())
        y = x[40:180] + bytes(rng.choice(list(b"ACGT"), 60).tolist()) + x[200:260]
        want = sorted((i, j) for i in range(len(x) - k + 1) for j in range(len(y) - k + 1) if x[i:i + k] == y[j:j + k])
        assert find_kmer_matches(x, y, k) == want
        assert len(want) > 0
    assert find_kmer_matches(b"ACGT", b"ACGTACGT", 20) == []       # shorter than k
    assert find_kmer_matches(TEXT, b"ACGTAC", 20) == []             # every known answer of the reference is band free


@pytest.mark.gpu
def test_host_mirror_replays_reference_tests():
    """The reference's own unit tests (src/bio_align_utils.rs:291-403) through the CUDA path and the host mirror."""
    from sawfish_b200.batch import AlignContext
    from sawfish_b200.bio_align_utils import AlignmentWeights, OverlapPairwiseAligner, PairwiseAligner, scoring_from_scores
    from sawfish_b200.simple_alignment import cigar_to_string
    ctx = AlignContext(0)
    w = AlignmentWeights(match_=1, mismatch=-3, gap_open=-1, gap_extend=-1)
    sc = scoring_from_scores(w.gap_open, w.gap_extend, w.match_, w.mismatch)
    makers = {"merge": lambda: PairwiseAligner.new_merge_aligner(ctx), "glocal": lambda: PairwiseAligner.new_glocal_aligner(ctx, w),
              "suffix": lambda: PairwiseAligner.new_pattern_suffix_aligner(ctx, sc, 20, 10),
              "prefix": lambda: PairwiseAligner.new_pattern_prefix_aligner(ctx, sc, 20, 10)}
    for flavour, pattern, score, aln in KNOWN:
        a = makers[flavour]()
        assert a.align(TEXT, pattern) == score
        assert a.band_free  # no shared 20-mer: rust-bio fills the whole matrix here too, so these answers pin K4 exactly
        if aln is not None:
            got = a.get_last_alignment()
            assert (got.ref_offset, cigar_to_string(got.cigar)) == aln
    ov = OverlapPairwiseAligner.new_aligner(ctx, w)  # test_overlap_aligner, :381-402
    assert ov.align(TEXT, b"ATAACCG") == 4
    got = ov.get_last_alignment()
    assert (got.ref_offset, cigar_to_string(got.cigar)) == (6, "4=3S")
    assert ov.align(TEXT, b"CCGAACT") == 4
    got = ov.get_last_alignment()
    assert (got.ref_offset, cigar_to_string(got.cigar)) == (0, "3S4=")


@pytest.mark.gpu
def test_k4_matches_oracle_on_random_pairs():
    from sawfish_b200.batch import AlignContext, PAIR_DTYPE
    ctx = AlignContext(0)
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGTN", np.uint8)
    for name, sc in scorings().items():
        seqs, rows, off = [], [], 0
        cases = []
        for q in range(60):
            L = int(rng.integers(2, 700 if q < 50 else 3000))
            base = acgt[rng.integers(0, 4, L)]
            x = base.copy()
            for _ in range(int(rng.integers(0, 6))):  # a few edits
                pos = int(rng.integers(0, len(x)))
                kind = int(rng.integers(0, 3))
                if kind == 0:
                    x[pos] = acgt[int(rng.integers(0, 5))]
                elif kind == 1 and len(x) > 40:
                    x = np.delete(x, slice(pos, pos + int(rng.integers(1, 30))))
                else:
                    x = np.insert(x, pos, acgt[rng.integers(0, 4, int(rng.integers(1, 30)))])
            if rng.random() < 0.5 and len(x) > 12:  # pattern covers only part of the text's core
                x = x[int(rng.integers(0, len(x) // 3)):len(x) - int(rng.integers(0, len(x) // 3))]
            y = np.concatenate([acgt[rng.integers(0, 4, int(rng.integers(0, 50)))], base, acgt[rng.integers(0, 4, int(rng.integers(0, 50)))]])
            cases.append((x.tobytes(), y.tobytes()))
            rows.append((off, off + len(x), len(x), len(y), 0, 0, 0, 0, 0))
            seqs += [x, y]
            off += len(x) + len(y)
        arena = np.concatenate(seqs)
        tab = np.array(rows, dtype=PAIR_DTYPE)
        scores, status, bounds, ops, ops_off = ctx.bio_align_batch(_to_gpu_scoring(sc), arena, tab, True)
        assert (status == 0).all()
        for i, (x, y) in enumerate(cases):
            s, xs, xe, ys, ye, o_ops = O.bio_custom_align(sc, x, y)
            assert scores[i] == s, (name, i, len(x), len(y))
            assert tuple(bounds[i]) == (xs, xe, ys, ye), (name, i)
            want = rle(o_ops)
            names = ("Match", "Subst", "Ins", "Del", "Xclip", "Yclip")
            got = [(names[int(w) & 7], int(w) >> 3) for w in ops[int(ops_off[i]):int(ops_off[i + 1])]]
            assert got == want, (name, i, got[:6], want[:6])


def _random_bio_cases(rng, n_small, n_large, max_large):
    acgt = np.frombuffer(b"ACGTN", np.uint8)
    cases = []
    for q in range(n_small + n_large):
        L = int(rng.integers(1, 700)) if q < n_small else int(rng.integers(700, max_large))
        base = acgt[rng.integers(0, 4, L)]
        x = base.copy()
        for _ in range(int(rng.integers(0, 6))):
            pos = int(rng.integers(0, len(x)))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                x[pos] = acgt[int(rng.integers(0, 5))]
            elif kind == 1 and len(x) > 40:
                x = np.delete(x, slice(pos, pos + int(rng.integers(1, 30))))
            else:
                x = np.insert(x, pos, acgt[rng.integers(0, 4, int(rng.integers(1, 30)))])
        if rng.random() < 0.5 and len(x) > 12:
            x = x[int(rng.integers(0, len(x) // 3)):len(x) - int(rng.integers(0, len(x) // 3))]
        if rng.random() < 0.15:  # unrelated pattern: the score is dominated by clips / gaps
            x = acgt[rng.integers(0, 4, int(rng.integers(1, 300)))]
        y = np.concatenate([acgt[rng.integers(0, 4, int(rng.integers(0, 50)))], base, acgt[rng.integers(0, 4, int(rng.integers(0, 50)))]])
        cases.append((x.tobytes(), y.tobytes()))
    return cases


def _pack_cases(cases):
    from sawfish_b200.batch import PAIR_DTYPE
    seqs, rows, off = [], [], 0
    for x, y in cases:
        rows.append((off, off + len(x), len(x), len(y), 0, 0, 0, 0, 0))
        seqs += [np.frombuffer(x, np.uint8), np.frombuffer(y, np.uint8)]
        off += len(x) + len(y)
    return np.concatenate(seqs), np.array(rows, dtype=PAIR_DTYPE)


@pytest.mark.gpu
def test_k5_score_only_matches_oracle():
    """sfc_bio_score_batch (K5, registers + shuffles, no traceback) == oracle score for every aligner flavour, on pairs
    that span one lane, one band, several bands of each height (R = 4, 8, 16) and band boundaries."""
    from sawfish_b200.batch import AlignContext
    ctx = AlignContext(0)
    rng = np.random.default_rng(23)
    extra = {
        "glocal_open": O.bio_scoring(-5, -1, 2, -4, yclip_prefix=0, yclip_suffix=0),
        "global": O.bio_scoring(-3, -1, 1, -2),
        "yprefix_only": O.bio_scoring(-1, -1, 1, -3, yclip_prefix=-2),
        "clips_paid": O.bio_scoring(-2, -1, 1, -3, xclip_prefix=-4, yclip_prefix=-3, yclip_suffix=-5),
    }
    for name, sc in {**scorings(), **extra}.items():
        cases = _random_bio_cases(rng, 70, 10, 2600)
        for mlen in (1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 383, 384, 385, 511, 512, 513, 640, 641, 1025):
            y = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(1, 900)))]
            x = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, mlen)]
            k = min(len(x), len(y))
            if k > 4 and rng.random() < 0.7:
                x[:k] = y[:k]
            cases.append((x.tobytes(), y.tobytes()))
        arena, tab = _pack_cases(cases)
        scores, status = ctx.bio_score_batch(_to_gpu_scoring(sc), arena, tab)
        assert (status == 0).all()
        for i, (x, y) in enumerate(cases):
            s = O.bio_custom_align(sc, x, y)[0]
            assert scores[i] == s, (name, i, len(x), len(y), int(scores[i]), s)


@pytest.mark.gpu
def test_k5_agrees_with_k4_on_merge_shaped_pairs():
    """Base contig x high-quality range of a near-identical contig (merge_pools.rs:214-272): K5 scores == K4 scores, and
    the merge decision (norm score >= 0.97) through the host mirror."""
    from sawfish_b200.batch import AlignContext
    from sawfish_b200.bio_align_utils import PairwiseAligner, test_for_duplicate_haplotypes_with_alignment
    ctx = AlignContext(0)
    rng = np.random.default_rng(29)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    cases = []
    for q in range(24):
        L = int(rng.integers(600, 6000))
        base = acgt[rng.integers(0, 4, L)]
        qy = base[int(0.1 * L):int(0.9 * L)].copy()
        n_edit = max(1, len(qy) // (200 if q % 3 else 20))  # every third query is too diverged to merge
        for _e in range(n_edit):
            pos = int(rng.integers(0, len(qy)))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                qy[pos] = acgt[int(rng.integers(0, 4))]
            elif kind == 1:
                qy = np.delete(qy, pos)
            else:
                qy = np.insert(qy, pos, acgt[int(rng.integers(0, 4))])
        cases.append((qy.tobytes(), base.tobytes()))
    ma = PairwiseAligner.new_merge_aligner(ctx)
    arena, tab = _pack_cases(cases)
    s5, st5 = ctx.bio_score_batch(ma.scoring, arena, tab)
    s4, st4, _b, _o, _oo = ctx.bio_align_batch(ma.scoring, arena, tab, False)
    assert (st5 == 0).all() and (st4 == 0).all()
    assert (s5 == s4).all()
    dup = test_for_duplicate_haplotypes_with_alignment(ctx, [(y, x) for x, y in cases])
    assert dup == [int(s) / len(x) >= 0.97 for s, (x, _y) in zip(s4, cases)]
    assert any(dup) and not all(dup)


def _k5_reduced_score(sc, x: bytes, y: bytes) -> int:
    """The recurrence K5 runs (sfc_k_bio_score.cuh), written out row by row: no y-prefix-clip candidate, no x-suffix clip,
    no traceback, final passes folded into one maximum.  Must equal the full oracle whenever K5 accepts the scoring."""
    go, ge, ma, mi = sc.gap_open, sc.gap_extend, sc.match, sc.mismatch
    xp, yp, ys = sc.xclip_prefix, sc.yclip_prefix, sc.yclip_suffix
    m, n, goe = len(x), len(y), sc.gap_open + sc.gap_extend
    xp_on = xp > O.BIO_MIN_SCORE // 2 if hasattr(O, "BIO_MIN_SCORE") else xp > -858993459 // 2
    s0 = [0] * (n + 1)
    sn0 = ys
    for j in range(1, n + 1):
        d0 = goe if j == 1 else max(go + ge * j, yp + goe)
        s0[j] = max(d0, yp)
        if j < n:
            sn0 = max(sn0, s0[j] + ys)
    s0[n] = max(s0[n], sn0)
    col0_i = lambda i: goe if i == 1 else max(go + ge * i, xp + goe)  # noqa: E731
    s_prev, i_prev = s0, [-858993459] * (n + 1)
    best = s0[n] + m * goe
    for i in range(1, m + 1):
        s_row, i_row = [0] * (n + 1), [0] * (n + 1)
        s_row[0], i_row[0] = max(col0_i(i), xp), col0_i(i)
        d, snmax = -858993459, s_row[0]
        for j in range(1, n + 1):
            ms = s_prev[j - 1] + (ma if x[i - 1] == y[j - 1] else mi)
            iv = max(i_prev[j] + ge, s_prev[j] + goe)
            d = max(d + ge, s_row[j - 1] + goe)
            s = max(ms, iv, d)
            if xp_on:
                s = max(s, xp + max(yp, go + ge * j))
            s_row[j], i_row[j] = s, iv
            snmax = max(snmax, s)
        best = max(best, max(s_row[n], snmax + ys) + (m - i) * goe)
        s_prev, i_prev = s_row, i_row
    return best


def test_k5_reduced_recurrence_equals_the_oracle_on_cpu():
    """The three simplifications K5 relies on (kernel header) hold for every scoring it accepts: x-suffix clip disabled,
    clip penalties <= 0.  Checked here on the CPU so the argument is pinned independently of the GPU run."""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    flavours = {
        "merge": O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0),
        "glocal": O.bio_scoring(-1, -1, 1, -3, yclip_prefix=0, yclip_suffix=0),
        "prefix": O.bio_scoring(-1, -1, 1, -3, yclip_suffix=0, xclip_prefix=0),
        "global": O.bio_scoring(-3, -1, 1, -2),
        "clips_paid": O.bio_scoring(-2, -1, 1, -3, xclip_prefix=-4, yclip_prefix=-3, yclip_suffix=-5),
        "ysuffix_only": O.bio_scoring(-2, -1, 1, -3, yclip_suffix=-1),
        "xprefix_only": O.bio_scoring(-2, -1, 1, -3, xclip_prefix=-1),
    }
    for name, sc in flavours.items():
        for _ in range(120):
            n = int(rng.integers(1, 36))
            y = acgt[rng.integers(0, 4, n)]
            if rng.random() < 0.6 and n > 3:
                a = int(rng.integers(0, n - 1))
                x = y[a:int(rng.integers(a + 1, n + 1))].copy()
                for _e in range(int(rng.integers(0, 3))):
                    x[int(rng.integers(0, len(x)))] = acgt[int(rng.integers(0, 4))]
                if rng.random() < 0.3:
                    x = np.insert(x, int(rng.integers(0, len(x) + 1)), acgt[rng.integers(0, 4, int(rng.integers(1, 5)))])
            else:
                x = acgt[rng.integers(0, 4, int(rng.integers(1, 36)))]
            xb, yb = x.tobytes(), y.tobytes()
            assert _k5_reduced_score(sc, xb, yb) == O.bio_custom_align(sc, xb, yb)[0], (name, xb, yb)


@pytest.mark.gpu
def test_k5_rejects_what_the_reference_asserts_on():
    """assert!(!text.is_empty()) / assert!(!pattern.is_empty()) (src/bio_align_utils.rs:177-178) become SFC_ERR_INVALID;
    so do pairs outside the arena and positive gap penalties.  A scoring with an enabled x-suffix clip is still served
    (by K4 behind the same entry point) with the oracle's score."""
    from sawfish_b200 import SfcError
    from sawfish_b200.batch import AlignContext, PAIR_DTYPE
    ctx = AlignContext(0)
    arena = np.frombuffer(b"ACGTACGTAC" + b"ACGTTCGTAC", np.uint8).copy()
    ok = np.array([(0, 10, 10, 10, 0, 0, 0, 0, 0)], dtype=PAIR_DTYPE)
    merge = _to_gpu_scoring(scorings()["merge"])
    scores, status = ctx.bio_score_batch(merge, arena, ok)
    assert status[0] == 0 and scores[0] == O.bio_custom_align(scorings()["merge"], b"ACGTACGTAC", b"ACGTTCGTAC")[0]
    for bad in [(0, 10, 0, 10, 0, 0, 0, 0, 0), (0, 10, 10, 0, 0, 0, 0, 0, 0), (0, 15, 10, 10, 0, 0, 0, 0, 0)]:
        with pytest.raises(SfcError) as ei:
            ctx.bio_score_batch(merge, arena, np.array([bad], dtype=PAIR_DTYPE))
        assert ei.value.code == -1
    pos_gap = _to_gpu_scoring(O.bio_scoring(1, -1, 1, -3))
    with pytest.raises(SfcError):
        ctx.bio_score_batch(pos_gap, arena, ok)
    suffix = scorings()["suffix"]  # x-suffix clip enabled: routed to K4
    s2, st2 = ctx.bio_score_batch(_to_gpu_scoring(suffix), arena, ok)
    assert st2[0] == 0 and s2[0] == O.bio_custom_align(suffix, b"ACGTACGTAC", b"ACGTTCGTAC")[0]
    ctx.close()
