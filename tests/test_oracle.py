"""CPU tests of the oracle itself: reference known answers, golden fixtures, DP optimality, internal consistency."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "wfa_golden.json")) as f:
        return json.load(f)


def test_reference_known_answer_translator(oracle):
    # reference src/wfa2_utils.rs:280-292
    roff, cig, hm = oracle.translate_cigar(b"DDDDMMMXXMMMMIIIMMMM"[::-1])
    assert hm and roff == 0 and oracle.cigar_to_string(cig) == "4S3=2X4=3D4="


def test_reference_known_answer_consensus(oracle):
    # reference src/wfa2_utils.rs:294-308
    for mode in (oracle.MODE_LITERAL, oracle.MODE_COMPACT):
        score, al, st = oracle.sawfish_align(oracle.CONSENSUS, b"AACTGTATAA", b"ACTTAT", mode=mode)
        assert score == 4 and al[0] == 1 and oracle.cigar_to_string(al[1]) == "3=1D3="
        assert st["wf_score"] == 5  # adjusted penalties x'=8, i'=5: one 1-base gap


def test_translator_quirks(oracle):
    # only the LAST run gets final_state: "...M DD II$" yields Ins(2) then drops II
    ops_fwd = b"MMMDDII"           # as process_wfa2_cigar_string iterates (reverse of the WFA string)
    roff, cig, hm = oracle.translate_cigar(ops_fwd[::-1])
    assert oracle.cigar_to_string(cig) == "3=2I" and roff == 0 and hm
    # no M/X at all -> None
    assert oracle.translate_cigar(b"IIDD")[2] is False
    # leading I before first match becomes ref_offset, leading D soft clip
    roff, cig, hm = oracle.translate_cigar(b"IIDDMM"[::-1][::-1][::-1])
    assert hm


def test_golden_vectors(oracle, golden):
    for ka in golden["reference_known_answers"]:
        if "text" in ka:
            score, al, _ = oracle.sawfish_align(oracle.CONSENSUS, ka["text"].encode(), ka["pattern"].encode())
            assert score == ka["score"] and al[0] == ka["ref_offset"] and oracle.cigar_to_string(al[1]) == ka["cigar"]
    pens = {"consensus": oracle.CONSENSUS, "large_indel": oracle.LARGE_INDEL}
    for v in golden["oracle_vectors"]:
        for mode in (oracle.MODE_LITERAL, oracle.MODE_COMPACT):
            r = oracle.wfa_align(pens[v["preset"]], v["pattern"].encode(), v["text"].encode(), *v["ends_free"], mode=mode)
            assert r.score == v["score"] and r.stats["wf_score"] == v["wf_score"] and r.ops.decode() == v["ops"]
            assert r.stats["k0"] == v["k0"]
        roff, cig, hm = oracle.translate_cigar(v["ops"].encode())
        assert (oracle.cigar_to_string(cig) if hm else None) == v["cigar"] and (not hm or roff == v["ref_offset"])


def _replay_penalty(pen_adj, metric, ops, p, t, ends):
    """Adjusted penalty of an op string, free ends discounted; also checks the ops spell out both sequences."""
    x, o1, e1, o2, e2 = pen_adj
    pbf, pef, tbf, tef = ends
    h = v = 0
    runs = []
    for ch in ops.decode():
        if runs and runs[-1][0] == ch:
            runs[-1][1] += 1
        else:
            runs.append([ch, 1])
    for ch, n in runs:
        for _ in range(n):
            if ch == "M":
                assert p[v] == t[h]
                h += 1; v += 1
            elif ch == "X":
                assert p[v] != t[h]
                h += 1; v += 1
            elif ch == "I":
                h += 1
            else:
                v += 1
    assert h == len(t) and v == len(p)

    def gap(n):
        if n <= 0:
            return 0
        return e1 * n if metric == 0 else min(o1 + e1 * n, o2 + e2 * n)
    if not any(ch in "MX" for ch, _ in runs):
        return 0  # degenerate all-gap alignment: begin/end free runs are not separable here
    total = 0
    # leading runs: 'I' run then 'D' run may be (partly) free
    i = 0
    lead = {"I": 0, "D": 0}
    while i < len(runs) and runs[i][0] in "ID":
        lead[runs[i][0]] += runs[i][1]
        i += 1
    j = len(runs)
    trail = {"I": 0, "D": 0}
    while j > i and runs[j - 1][0] in "ID":
        trail[runs[j - 1][0]] += runs[j - 1][1]
        j -= 1
    total += gap(lead["I"] - min(lead["I"], tbf)) + gap(lead["D"] - min(lead["D"], pbf))
    total += gap(trail["I"] - min(trail["I"], tef)) + gap(trail["D"] - min(trail["D"], pef))
    for ch, n in runs[i:j]:
        if ch == "X":
            total += x * n
        elif ch in "ID":
            total += gap(n)
    return total


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_oracle_optimal_and_consistent(oracle, seed):
    """wf_score == independent full-DP optimum; literal == compact; ops replay to <= that penalty bound."""
    from sawfish_b200 import synth
    arena, pairs = synth.random_pairs(400, seed=seed, max_len=80)
    for pen in (oracle.CONSENSUS, oracle.LARGE_INDEL):
        adj = oracle.adjusted(pen)
        for pr in pairs:
            p = arena[int(pr["pattern_off"]):int(pr["pattern_off"]) + int(pr["pattern_len"])].tobytes()
            t = arena[int(pr["text_off"]):int(pr["text_off"]) + int(pr["text_len"])].tobytes()
            if pr["flags"] & 1:
                p, t = p[::-1], t[::-1]
            ends = [int(pr[k]) for k in ("pattern_begin_free", "pattern_end_free", "text_begin_free", "text_end_free")]
            lit = oracle.wfa_align(pen, p, t, *ends, mode=oracle.MODE_LITERAL)   # internally cross-checks both backtraces
            cmp_ = oracle.wfa_align(pen, p, t, *ends, mode=oracle.MODE_COMPACT)
            assert lit.ops == cmp_.ops and lit.score == cmp_.score and lit.stats["wf_score"] == cmp_.stats["wf_score"]
            assert oracle.dp_min_penalty(pen, p, t, *ends) == lit.stats["wf_score"]
            # the emitted ops are a valid alignment whose (free-end-discounted) penalty is the optimum
            assert _replay_penalty(adj, pen.metric, lit.ops, p, t, ends) <= lit.stats["wf_score"]


def test_score_formula_free_ends(oracle):
    # perfect match inside a longer text: all text overhang is free -> score = pattern length, wf_score 0
    t = b"GGGGG" + b"ACGTACGTTGCA" + b"CCCCC"
    p = b"ACGTACGTTGCA"
    r = oracle.wfa_align(oracle.CONSENSUS, p, t, 0, 0, 5, 5)
    assert r.stats["wf_score"] == 0 and r.score == len(p) and r.stats["k0"] == 5
    # end-to-end on the same pair must pay for the overhang
    r2 = oracle.wfa_align(oracle.CONSENSUS, p, t, 0, 0, 0, 0)
    assert r2.stats["wf_score"] == 10 * 5 and r2.score == len(p) - 2 * 10


def test_batch_driver_matches_single(oracle):
    from sawfish_b200 import synth
    arena, pairs, _ = synth.config2b_score_sv_flanks(20, seed=5)
    scores, status, stats, ops, off = oracle.align_batch(oracle.CONSENSUS, arena, pairs.astype(oracle.PAIR_DTYPE), True,
                                                         oracle.MODE_COMPACT, 2)
    assert (status == 0).all()
    for i in (0, 7, 39):
        pr = pairs[i]
        p = arena[int(pr["pattern_off"]):int(pr["pattern_off"]) + int(pr["pattern_len"])].tobytes()[::-1]
        t = arena[int(pr["text_off"]):int(pr["text_off"]) + int(pr["text_len"])].tobytes()[::-1]
        c = int(pr["pattern_begin_free"])
        r = oracle.wfa_align(oracle.CONSENSUS, p, t, c, c, c, c)
        assert r.score == scores[i] and r.ops == ops[int(off[i]):int(off[i]) + int(stats["n_ops"][i])].tobytes()


def test_low_complexity_literal_equals_compact(oracle):
    from sawfish_b200 import synth
    arena, pairs = synth.low_complexity_pairs(40, seed=5, len_lo=100, len_hi=500)
    for pen in (oracle.CONSENSUS, oracle.LARGE_INDEL):
        for pr in pairs:
            p = arena[int(pr["pattern_off"]):int(pr["pattern_off"]) + int(pr["pattern_len"])].tobytes()
            t = arena[int(pr["text_off"]):int(pr["text_off"]) + int(pr["text_len"])].tobytes()
            if pr["flags"] & 1:
                p, t = p[::-1], t[::-1]
            ends = [int(pr[k]) for k in ("pattern_begin_free", "pattern_end_free", "text_begin_free", "text_end_free")]
            lit = oracle.wfa_align(pen, p, t, *ends, mode=oracle.MODE_LITERAL)   # -4 if the two backtraces disagree
            assert oracle.dp_min_penalty(pen, p, t, *ends) == lit.stats["wf_score"]
