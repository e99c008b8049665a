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
