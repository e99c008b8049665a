"""GPU parity tests proper: the CUDA path (through the C ABI) vs the CPU oracle, bit-exact.

Run on the B200 box: python -m pytest tests -m gpu
"""
import numpy as np
import pytest

from helpers import compare_with_oracle

pytestmark = pytest.mark.gpu


def _pens():
    from sawfish_b200.batch import CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES
    return CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES


def test_native_library_loaded(ctx):
    from sawfish_b200 import _lib
    assert _lib.lib().sfc_device_count() >= 1


def test_known_answer_consensus_aligner(ctx):
    # reference src/wfa2_utils.rs:294-308
    from sawfish_b200.simple_alignment import cigar_to_string
    from sawfish_b200.wfa2_utils import PairwiseAligner
    al = PairwiseAligner.new_consensus_aligner(ctx)
    al.set_text(b"AACTGTATAA")
    score, a = al.align(b"ACTTAT")
    assert score == 4
    assert a is not None and a.ref_offset == 1 and cigar_to_string(a.cigar) == "3=1D3="


def test_pairwise_aligner_surface(ctx, oracle):
    from sawfish_b200.simple_alignment import cigar_to_u32
    from sawfish_b200.wfa2_utils import PairwiseAligner
    rng = np.random.default_rng(7)
    text = bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 400)])
    pat = text[50:180] + b"ACGTTGCA" * 6 + text[180:350]
    for ctor, pen in ((PairwiseAligner.new_consensus_aligner, oracle.CONSENSUS),
                      (lambda c: PairwiseAligner.new_large_indel_aligner(False, c), oracle.LARGE_INDEL)):
        al = ctor(ctx)
        al.set_text(text)
        assert al.text() == text[::-1]
        for fn, kw in ((al.align, {}), (lambda p: al.align_clip_frac(p, 0.3, 0.1), dict(text_clip_frac=0.3, pattern_clip_frac=0.1)),
                       (al.align_end_to_end, dict(end_to_end=True))):
            score, a = fn(pat)
            o_score, o_al, _ = oracle.sawfish_align(pen, text, pat, **kw)
            assert score == o_score
            assert (a is None) == (o_al is None)
            if a is not None:
                assert a.ref_offset == o_al[0] and cigar_to_u32(a.cigar) == o_al[1]
    with pytest.raises(AssertionError):
        PairwiseAligner.new_consensus_aligner(ctx).align(b"ACGT")  # text not set
    with pytest.raises(AssertionError):
        al.align(b"")


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("which", ["linear_score", "linear_cigar", "affine_cigar", "affine_score"])
def test_random_small_pairs(ctx, oracle, seed, which):
    from sawfish_b200 import synth
    lin, aff = _pens()
    arena, pairs = synth.random_pairs(1500, seed=seed * 101 + len(which), max_len=150)
    pen = lin if which.startswith("linear") else aff
    compare_with_oracle(oracle, ctx, pen, arena, pairs, want_cigar=which.endswith("cigar"), mode=oracle.MODE_LITERAL)


@pytest.mark.parametrize("which", ["linear_score", "linear_cigar", "affine_cigar"])
def test_low_complexity_mid_size_pairs(ctx, oracle, which):
    """Tandem repeats / homopolymers / 2-letter sequences / N runs, 300-3000 bp: many co-optimal alignments, so CIGAR
    parity here exercises every tie-break."""
    from sawfish_b200 import synth
    lin, aff = _pens()
    arena, pairs = synth.low_complexity_pairs(160, seed=len(which) * 7 + 1)
    pen = lin if which.startswith("linear") else aff
    compare_with_oracle(oracle, ctx, pen, arena, pairs, want_cigar=which.endswith("cigar"))


def test_score_sv_shaped_batch(ctx, oracle):
    # config 2b: 100-500 bp read flank vs ref/alt allele flanks, gap-linear, score only (K1)
    from sawfish_b200 import synth
    lin, _ = _pens()
    arena, pairs, _ = synth.config2b_score_sv_flanks(3000)
    compare_with_oracle(oracle, ctx, lin, arena, pairs, want_cigar=False)
    st = ctx.stats()
    assert st["n_score_kernel"] == len(pairs)
    compare_with_oracle(oracle, ctx, lin, arena[:], pairs[:400], want_cigar=True)


def test_refine_shaped_batch(ctx, oracle):
    # config 3: contig vs reference window incl. 2-region alt-hap with the 100xN spacer, gap-affine-2p + CIGAR
    from sawfish_b200 import synth
    _, aff = _pens()
    arena, pairs = synth.config3_refine_contigs(24, two_region_frac=0.3)
    compare_with_oracle(oracle, ctx, aff, arena, pairs, want_cigar=True)


def test_config2_reads_vs_haplotypes(ctx, oracle):
    from sawfish_b200 import synth
    lin, aff = _pens()
    arena, pairs, _ = synth.config2_read_vs_haplotypes(8, len_lo=2000, len_hi=6000, sv_hi=1500)
    compare_with_oracle(oracle, ctx, aff, arena, pairs, want_cigar=True)
    compare_with_oracle(oracle, ctx, lin, arena, pairs, want_cigar=False)


def test_edge_cases(ctx, oracle):
    from sawfish_b200.batch import PAIR_DTYPE
    lin, aff = _pens()
    seqs = [b"A", b"C", b"N", b"NNNNNNNN", b"ACGT" * 10, b"A" * 40, b"AC" * 30, b"ACGTN" * 8, b"T" * 7 + b"G" + b"T" * 9]
    arena = np.frombuffer(b"".join(seqs), np.uint8).copy()
    offs = np.cumsum([0] + [len(s) for s in seqs])
    rows = []
    for i, p in enumerate(seqs):
        for j, t in enumerate(seqs):
            c = min(int(len(p) * 0.4), int(len(t) * 0.4))
            rows.append((offs[i], offs[j], len(p), len(t), c, c, c, c, 1))
            rows.append((offs[i], offs[j], len(p), len(t), 0, 0, 0, 0, 0))
            rows.append((offs[i], offs[j], len(p), len(t), len(p), 0, 0, len(t), 0))
    pairs = np.array(rows, dtype=PAIR_DTYPE)
    for pen in (lin, aff):
        for wc in (False, True):
            compare_with_oracle(oracle, ctx, pen, arena, pairs, want_cigar=wc, mode=oracle.MODE_LITERAL)


def test_invalid_inputs_fail_loudly(ctx):
    from sawfish_b200.batch import PAIR_DTYPE, CONSENSUS_PENALTIES
    from sawfish_b200 import SfcError
    arena = np.frombuffer(b"ACGTACGTAC", np.uint8).copy()
    ok = np.array([(0, 4, 4, 6, 1, 1, 1, 1, 1)], dtype=PAIR_DTYPE)
    for bad in [(0, 4, 0, 6, 0, 0, 0, 0, 1), (0, 4, 4, 60, 0, 0, 0, 0, 1), (0, 4, 4, 6, 5, 0, 0, 0, 1)]:
        with pytest.raises(SfcError):
            ctx.align_batch(CONSENSUS_PENALTIES, arena, np.array([bad], dtype=PAIR_DTYPE), False)
    bad_sym = np.frombuffer(b"ACGTAC*TAC", np.uint8).copy()
    with pytest.raises(SfcError):
        ctx.align_batch(CONSENSUS_PENALTIES, bad_sym, ok, False)
    s, st, _, _ = ctx.align_batch(CONSENSUS_PENALTIES, arena, ok, False)
    assert st[0] == 0


def test_single_process_multi_gpu_sharding(ctx):
    """sfc_multi_*: shard by SV group over every visible GPU, gather on the host; identical to one-GPU results."""
    from sawfish_b200 import _lib, synth
    from sawfish_b200.batch import MultiGpuContext
    lin, aff = _pens()
    ndev = _lib.lib().sfc_device_count()
    devices = list(range(min(ndev, 8)))
    multi = MultiGpuContext(devices)
    try:
        arena, pairs, _ = synth.config2b_score_sv_flanks(4000, seed=21)
        groups = np.arange(len(pairs), dtype=np.uint32) // 8
        s1, st1, _, _ = ctx.align_batch(lin, arena, pairs, False)
        s2, st2, _, _ = multi.align_batch(lin, arena, pairs, False, group_of_pair=groups)
        assert (st2 == 0).all() and (s1 == s2).all()
        arena, pairs = synth.config3_refine_contigs(40, seed=22)
        s1, st1, o1, f1 = ctx.align_batch(aff, arena, pairs, True)
        s2, st2, o2, f2 = multi.align_batch(aff, arena, pairs, True)
        assert (st2 == 0).all() and (s1 == s2).all() and (f1 == f2).all() and (o1 == o2).all()
        if len(devices) > 1:
            assert (multi.last_kernel_ms() > 0).all()  # every GPU got work
    finally:
        multi.close()


def test_asymmetric_lengths_and_clips(ctx, oracle):
    """Short pattern inside a long text and vice versa, zero / maximal / one-sided ends-free allowances."""
    from sawfish_b200 import synth
    from sawfish_b200.batch import PAIR_DTYPE
    lin, aff = _pens()
    rng = np.random.default_rng(12)
    long_seq = synth.random_seq(rng, 6000)
    short = synth.add_errors(rng, long_seq[2500:2740], 0.01)
    mid = synth.add_errors(rng, np.concatenate([long_seq[1000:1800], long_seq[1900:2600]]), 0.002)
    arena = np.concatenate([long_seq, short, mid])
    oL, oS, oM = 0, len(long_seq), len(long_seq) + len(short)
    nL, nS, nM = len(long_seq), len(short), len(mid)
    rows = []
    for (po, pn, to, tn) in [(oS, nS, oL, nL), (oL, nL, oS, nS), (oM, nM, oL, nL), (oL, nL, oM, nM), (oS, nS, oM, nM)]:
        for ends in [(0, 0, 0, 0), (0, 0, tn, tn), (pn, pn, 0, 0), (pn // 2, 0, 0, tn // 2), (0, pn // 3, tn // 3, 0),
                     (min(pn, tn) * 2 // 5,) * 4]:
            ends = (min(ends[0], pn), min(ends[1], pn), min(ends[2], tn), min(ends[3], tn))
            for rev in (0, 1):
                rows.append((po, to, pn, tn) + ends + (rev,))
    pairs = np.array(rows, dtype=PAIR_DTYPE)
    compare_with_oracle(oracle, ctx, lin, arena, pairs, want_cigar=False)
    compare_with_oracle(oracle, ctx, lin, arena, pairs, want_cigar=True)
    compare_with_oracle(oracle, ctx, aff, arena, pairs, want_cigar=True)


def test_ops_pool_growth_path(ctx, oracle):
    """Enough CIGAR runs in one batch to overflow the initial device ops pool (1 Mi entries): the pairs that did not
    fit are re-run with a larger pool (stats.retries > 0) and every result still matches the oracle."""
    from sawfish_b200 import synth
    from sawfish_b200.batch import PAIR_DTYPE, sawfish_clip
    _, aff = _pens()
    rng = np.random.default_rng(99)
    chunks, rows, off = [], [], 0
    for _ in range(9000):
        t = synth.random_seq(rng, 900)
        p = synth.add_errors(rng, t, 0.09)
        chunks += [t, p]
        c = sawfish_clip(len(p), len(t))
        rows.append((off + len(t), off, len(p), len(t), c, c, c, c, 1))
        off += len(t) + len(p)
    arena, pairs = np.concatenate(chunks), np.array(rows, dtype=PAIR_DTYPE)
    compare_with_oracle(oracle, ctx, aff, arena, pairs, want_cigar=True)
    assert ctx.stats()["retries"] >= 1


@pytest.mark.parametrize("preset", ["large_indel_cigar", "consensus_score", "consensus_cigar"])
def test_config2_at_bench_size(ctx, oracle, preset):
    """The configuration bench.py times (BASELINE configs[1]: 2-20 kb reads x 2 haplotypes, SV up to 5 kb) against the
    oracle at full size: scores and run-length ops bit-exact (reference src/wfa2_utils.rs:209-219)."""
    from sawfish_b200 import synth
    lin, aff = _pens()
    arena, pairs, _ = synth.config2_read_vs_haplotypes(64, seed=0x5AF15 + 2, len_lo=2000, len_hi=20000, sv_lo=35, sv_hi=5000)
    assert int(pairs["text_len"].max()) > 15000
    pen, wc = {"large_indel_cigar": (aff, True), "consensus_score": (lin, False), "consensus_cigar": (lin, True)}[preset]
    compare_with_oracle(oracle, ctx, pen, arena, pairs, want_cigar=wc)
    assert ctx.stats()["retries"] == 0


def test_config2_bench_size_int16_rings(ctx, oracle, monkeypatch):
    from sawfish_b200 import synth
    _, aff = _pens()
    monkeypatch.setenv("SFC_K2_NARROW", "1")
    arena, pairs, _ = synth.config2_read_vs_haplotypes(48, seed=4242, len_lo=2000, len_hi=20000, sv_lo=35, sv_hi=5000)
    compare_with_oracle(oracle, ctx, aff, arena, pairs, want_cigar=True)


def test_config5_stress_at_size(ctx, oracle):
    """BASELINE configs[4]: 10 kb+ insertion / duplication haplotypes at 1-5 % divergence, full CIGAR parity at a size
    where the wavefronts are tens of thousands of diagonals wide and thousands of scores deep."""
    from sawfish_b200 import synth
    _, aff = _pens()
    arena, pairs = synth.config5_stress(3, seed=0x5AF15 + 5, len_lo=10000, len_hi=16000)
    assert int(pairs["pattern_len"].min()) >= 10000
    compare_with_oracle(oracle, ctx, aff, arena, pairs, want_cigar=True)


def test_arena_may_be_reused_as_soon_as_upload_returns(ctx, oracle):
    """include/sawfish_cuda.h, buffer lifetime: sfc_batch_upload has finished reading the caller's (pinned) arena and pair
    table when it returns, so a pipelined submitter may refill both right away."""
    import torch
    from sawfish_b200 import synth
    lin, _ = _pens()
    arena, pairs, _ = synth.config2b_score_sv_flanks(20000, seed=31)
    want, st, _, _ = ctx.align_batch(lin, arena, pairs, False)
    assert (st == 0).all()
    a_pin = torch.from_numpy(arena.copy()).pin_memory()
    p_pin = torch.from_numpy(pairs.view(np.uint8).copy()).pin_memory()
    for _ in range(3):
        a_pin.numpy()[:] = arena
        p_pin.numpy()[:] = pairs.view(np.uint8)
        ctx.upload_ptr(a_pin.data_ptr(), a_pin.numel(), p_pin.data_ptr(), len(pairs))
        a_pin.numpy()[:] = ord("N")      # overwrite immediately: the library must not read these buffers any more
        p_pin.numpy()[:] = 0
        ctx.run(lin, False)
        got, st, _, _ = ctx.download()
        assert (st == 0).all() and (got == want).all()


def test_matching_debug_view(ctx):
    """PairwiseAligner::matching / print_pairwise_alignment (reference src/wfa2_utils.rs:253-273): the three display lines
    of the last alignment spell out both (reversed) sequences and mark exactly the matching columns."""
    import io
    from sawfish_b200.wfa2_utils import PairwiseAligner, print_pairwise_alignment
    al = PairwiseAligner.new_consensus_aligner(ctx)
    text, pat = b"AACTGTATAA", b"ACTTAT"
    al.set_text(text)
    al.align(pat)
    l1, l2, l3 = al.matching(al.text(), pat[::-1])
    assert len(l1) == len(l2) == len(l3)
    assert l1.replace("-", "").encode() == pat[::-1] and l3.replace("-", "").encode() == text[::-1]
    for a, m, b in zip(l1, l2, l3):
        assert (m == "|") == (a == b and a != "-")
    assert l2.count("|") == 6
    out = io.StringIO()
    print_pairwise_alignment(al, pat, file=out)
    rows = out.getvalue().splitlines()
    assert rows[1].replace("-", "").encode() == text and rows[3].replace("-", "").encode() == pat and "|" in rows[2]


# (metric, match, mismatch, gap_open1, gap_ext1, gap_open2, gap_ext2): the ABI takes any such set, sawfish uses two of them.
# Linear: mismatch cheaper / dearer than an indel, with and without a match bonus (the ring depth, which source row is the
# older one and which scores exist all change); affine-2p: other slopes and a second piece that never / always wins.
_OTHER_PENALTIES = [
    (0, 0, 4, 0, 2, 0, 0), (0, 0, 1, 0, 3, 0, 0), (0, -2, 5, 0, 1, 0, 0), (0, -1, 2, 0, 6, 0, 0), (0, 0, 7, 0, 7, 0, 0),
    (1, 0, 4, 6, 2, 24, 1), (1, -1, 2, 3, 1, 20, 0), (1, 0, 5, 2, 3, 2, 3), (1, -2, 6, 1, 4, 40, 1),
]


@pytest.mark.parametrize("pen_t", _OTHER_PENALTIES, ids=lambda p: "-".join(str(abs(v)) for v in p))
def test_penalty_sets_other_than_sawfish_presets(ctx, oracle, pen_t):
    from sawfish_b200 import synth
    from sawfish_b200.batch import Penalties
    pen = Penalties(*pen_t)
    arena, pairs = synth.random_pairs(600, seed=sum(abs(v) * (i + 1) for i, v in enumerate(pen_t)), max_len=300)
    compare_with_oracle(oracle, ctx, pen, arena, pairs, want_cigar=False, mode=oracle.MODE_LITERAL)
    compare_with_oracle(oracle, ctx, pen, arena, pairs[:200], want_cigar=True, mode=oracle.MODE_LITERAL)
    arena, pairs = synth.low_complexity_pairs(40, seed=pen_t[2] * 13 + pen_t[4], len_lo=300, len_hi=1500)
    compare_with_oracle(oracle, ctx, pen, arena, pairs, want_cigar=False)
    compare_with_oracle(oracle, ctx, pen, arena, pairs, want_cigar=True)
