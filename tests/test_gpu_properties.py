"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (no oracle needed at these sizes):
valid CIGARs whose own score equals the reported score, determinism, batch-order invariance, K1 == K2 scores,
identity alignments, and recovery of the planted truth."""
import numpy as np
import pytest

from helpers import replay_rle, sw_score_of_runs

pytestmark = pytest.mark.gpu


def _seqs(arena, pr):
    p = arena[int(pr["pattern_off"]):int(pr["pattern_off"]) + int(pr["pattern_len"])].tobytes()
    t = arena[int(pr["text_off"]):int(pr["text_off"]) + int(pr["text_len"])].tobytes()
    if pr["flags"] & 1:
        p, t = p[::-1], t[::-1]
    return p, t


def _ends(pr):
    return [int(pr[k]) for k in ("pattern_begin_free", "pattern_end_free", "text_begin_free", "text_end_free")]


@pytest.mark.parametrize("preset", ["large_indel", "consensus"])
def test_config2_full_size_cigars_are_valid_and_score_consistent(ctx, preset):
    from sawfish_b200 import synth
    from sawfish_b200.batch import CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES
    pen = LARGE_INDEL_PENALTIES if preset == "large_indel" else CONSENSUS_PENALTIES
    n_reads = 48 if preset == "large_indel" else 12
    arena, pairs, truth = synth.config2_read_vs_haplotypes(n_reads, seed=4242, len_lo=2000, len_hi=20000)
    scores, status, ops, off = ctx.align_batch(pen, arena, pairs, True)
    assert (status == 0).all()
    for i, pr in enumerate(pairs):
        p, t = _seqs(arena, pr)
        runs = replay_rle(ops[int(off[i]):int(off[i + 1])], p, t)
        sc = sw_score_of_runs(runs, pen, _ends(pr))
        assert sc == scores[i], f"pair {i}: score of the emitted CIGAR {sc} != reported {scores[i]}"
    # planted truth: the haplotype the read was sampled from scores higher
    for r in range(n_reads):
        assert scores[2 * r + truth[r]] > scores[2 * r + 1 - truth[r]]


def test_determinism_and_batch_order_invariance(ctx):
    from sawfish_b200 import synth
    from sawfish_b200.batch import LARGE_INDEL_PENALTIES
    arena, pairs = synth.config3_refine_contigs(96, seed=31)
    s1, st1, o1, f1 = ctx.align_batch(LARGE_INDEL_PENALTIES, arena, pairs, True)
    s2, st2, o2, f2 = ctx.align_batch(LARGE_INDEL_PENALTIES, arena, pairs, True)
    assert (st1 == 0).all() and (s1 == s2).all() and (f1 == f2).all() and (o1 == o2).all()
    perm = np.random.default_rng(5).permutation(len(pairs))
    s3, st3, o3, f3 = ctx.align_batch(LARGE_INDEL_PENALTIES, arena, np.ascontiguousarray(pairs[perm]), True)
    assert (s3 == s1[perm]).all()
    for j, i in enumerate(perm):
        assert (o3[int(f3[j]):int(f3[j + 1])] == o1[int(f1[i]):int(f1[i + 1])]).all()


def test_score_kernel_equals_backtrace_kernel_at_scale(ctx):
    """K1 (score only) and K2-linear (with CIGAR) are independent kernels: their scores must agree on 120k pairs."""
    from sawfish_b200 import synth
    from sawfish_b200.batch import CONSENSUS_PENALTIES
    arena, pairs, _ = synth.config2b_score_sv_flanks(60000, seed=77)
    s1, st1, _, _ = ctx.align_batch(CONSENSUS_PENALTIES, arena, pairs, False)
    assert ctx.stats()["n_score_kernel"] == len(pairs)
    s2, st2, ops, off = ctx.align_batch(CONSENSUS_PENALTIES, arena, pairs, True)
    assert ctx.stats()["n_align_kernel"] == len(pairs)
    assert (st1 == 0).all() and (st2 == 0).all()
    assert (s1 == s2).all(), f"{int((s1 != s2).sum())} score differences between K1 and K2"


def test_identity_alignments(ctx):
    from sawfish_b200 import synth
    from sawfish_b200.batch import CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES, PAIR_DTYPE, sawfish_clip
    rng = np.random.default_rng(1)
    lens = [1, 7, 8, 9, 255, 256, 257, 500, 4095, 20000, 33000]
    seqs = [synth.random_seq(rng, n) for n in lens]
    arena = np.concatenate(seqs)
    offs = np.cumsum([0] + lens)
    rows = [(offs[i], offs[i], n, n, sawfish_clip(n, n), sawfish_clip(n, n), sawfish_clip(n, n), sawfish_clip(n, n), 1)
            for i, n in enumerate(lens)]
    pairs = np.array(rows, dtype=PAIR_DTYPE)
    for pen in (CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES):
        scores, status, ops, off = ctx.align_batch(pen, arena, pairs, True)
        assert (status == 0).all() and (scores == np.array(lens)).all()
        for i, n in enumerate(lens):
            assert list(ops[int(off[i]):int(off[i + 1])]) == [(n << 2) | 0]
        s0, st0, _, _ = ctx.align_batch(pen, arena, pairs, False)
        assert (s0 == scores).all()
