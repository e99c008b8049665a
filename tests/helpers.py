"""Shared comparison helpers for the parity tests (oracle = checker, never the thing under test)."""
import numpy as np


def oracle_batch(O, pen_gpu, arena, pairs, want_ops=True, mode=None, threads=0):
    pen = O.Penalties(pen_gpu.metric, pen_gpu.match, pen_gpu.mismatch, pen_gpu.gap_open1, pen_gpu.gap_ext1,
                      pen_gpu.gap_open2, pen_gpu.gap_ext2)
    opairs = pairs.astype(O.PAIR_DTYPE)
    mode = O.MODE_COMPACT if mode is None else mode
    return O.align_batch(pen, arena, opairs, want_ops=want_ops, mode=mode, n_threads=threads)


def rle_of_bytes(ops: bytes):
    """Run-length encode a WFA2 op string into the ABI's (len << 2 | code) words."""
    code = {ord("M"): 0, ord("X"): 1, ord("I"): 2, ord("D"): 3}
    a = np.frombuffer(ops, np.uint8)
    if len(a) == 0:
        return np.zeros(0, np.uint32)
    brk = np.flatnonzero(np.diff(a)) + 1
    starts = np.concatenate([[0], brk])
    lens = np.diff(np.concatenate([starts, [len(a)]]))
    return np.array([(int(l) << 2) | code[int(a[s])] for s, l in zip(starts, lens)], np.uint32)


def compare_with_oracle(O, ctx, pen, arena, pairs, want_cigar, mode=None):
    """Run the CUDA path and the oracle on the same batch; assert bit-exact scores (and run-length ops)."""
    scores, status, ops, ops_off = ctx.align_batch(pen, arena, pairs, want_cigar)
    o_scores, o_status, o_stats, o_ops, o_off = oracle_batch(O, pen, arena, pairs, want_ops=want_cigar, mode=mode)
    assert (o_status == 0).all(), f"oracle failed on {np.flatnonzero(o_status != 0)[:5]}"
    bad = np.flatnonzero(status != 0)
    assert len(bad) == 0, f"GPU status != 0 for pairs {bad[:5]}: {status[bad[:5]]}"
    mism = np.flatnonzero(scores != o_scores)
    assert len(mism) == 0, (f"{len(mism)} score mismatches, first pair {mism[0]}: gpu {scores[mism[0]]} "
                            f"oracle {o_scores[mism[0]]} pair={pairs[mism[0]]}")
    if want_cigar:
        for i in range(len(pairs)):
            want = rle_of_bytes(o_ops[int(o_off[i]):int(o_off[i]) + int(o_stats["n_ops"][i])].tobytes())
            got = ops[int(ops_off[i]):int(ops_off[i + 1])]
            assert len(want) == len(got) and (want == got).all(), f"pair {i}: ops differ\n gpu {got[:12]}\n ora {want[:12]}"
    return scores, o_stats


def replay_rle(ops_rle, p, t):
    """Replay run-length ops (ABI codes) over pattern p / text t (as aligned, i.e. already reversed if the pair is).
    Checks that the ops spell out both sequences and returns the list of (code, len) runs."""
    h = v = 0
    runs = []
    p = np.frombuffer(p, np.uint8)
    t = np.frombuffer(t, np.uint8)
    for w in ops_rle:
        code, n = int(w) & 3, int(w) >> 2
        runs.append((code, n))
        if code == 0:
            assert (p[v:v + n] == t[h:h + n]).all(), "M run over unequal bases"
            h += n; v += n
        elif code == 1:
            assert (p[v:v + n] != t[h:h + n]).all(), "X run over equal bases"
            h += n; v += n
        elif code == 2:
            h += n
        else:
            v += n
    assert h == len(t) and v == len(p), "ops do not consume both sequences"
    return runs


def sw_score_of_runs(runs, pen, ends):
    """SW-scale score of the non-free part of an alignment under sawfish's penalties (match +1): the value
    WFA2 reports. Leading/trailing gap runs are free up to the ends-free allowances."""
    m, x, o1, e1, o2, e2 = -pen.match, pen.mismatch, pen.gap_open1, pen.gap_ext1, pen.gap_open2, pen.gap_ext2
    pbf, pef, tbf, tef = ends

    def gap(n):
        if n <= 0:
            return 0
        return e1 * n if pen.metric == 0 else min(o1 + e1 * n, o2 + e2 * n)
    if not any(c in (0, 1) for c, _ in runs):
        return None
    i = 0
    lead = {2: 0, 3: 0}
    while runs[i][0] in (2, 3):
        lead[runs[i][0]] += runs[i][1]; i += 1
    j = len(runs)
    trail = {2: 0, 3: 0}
    while runs[j - 1][0] in (2, 3):
        trail[runs[j - 1][0]] += runs[j - 1][1]; j -= 1
    score = -(gap(lead[2] - min(lead[2], tbf)) + gap(lead[3] - min(lead[3], pbf)) + gap(trail[2] - min(trail[2], tef)) +
              gap(trail[3] - min(trail[3], pef)))
    for c, n in runs[i:j]:
        score += m * n if c == 0 else (-x * n if c == 1 else -gap(n))
    return score
