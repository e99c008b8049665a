#!/usr/bin/env python
"""K4 (rust-bio style merge aligner) on merge-shaped pairs: base contig 0.6-10 kb x query high-quality range of a
near-identical contig (reference src/joint_call/merge_haplotypes/single_region/merge_pools.rs:214-272).
Prints one JSON line: GPU alignments/s and GCUPS (cells = |x| * |y|), CPU oracle (one thread) on a bounded sample, parity
of scores on that sample.  usage: python tests/tools/bench_merge.py [n_pairs] [cpu_seconds]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (CPU baseline leg only)
from sawfish_b200.batch import AlignContext, PAIR_DTYPE  # noqa: E402
from sawfish_b200.bio_align_utils import PairwiseAligner  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 592
cpu_s = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
rng = np.random.default_rng(0x5AF15 + 30)
acgt = np.frombuffer(b"ACGT", np.uint8)
seqs, rows, off, cases = [], [], 0, []
for _ in range(n_pairs):
    L = int(rng.integers(600, 10001))
    base = acgt[rng.integers(0, 4, L)]
    q = base[int(0.1 * L):int(0.9 * L)].copy()
    for _e in range(max(1, len(q) // 200)):  # ~0.5 % edits
        pos = int(rng.integers(0, len(q)))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            q[pos] = acgt[int(rng.integers(0, 4))]
        elif kind == 1:
            q = np.delete(q, pos)
        else:
            q = np.insert(q, pos, acgt[int(rng.integers(0, 4))])
    rows.append((off, off + len(q), len(q), len(base), 0, 0, 0, 0, 0))  # x = query (pattern), y = base contig (text)
    seqs += [q, base]
    cases.append((q.tobytes(), base.tobytes()))
    off += len(q) + len(base)
arena = np.concatenate(seqs)
tab = np.array(rows, dtype=PAIR_DTYPE)
ctx = AlignContext(0)
sc = PairwiseAligner.new_merge_aligner(ctx).scoring
ctx.bio_align_batch(sc, arena, tab[:8], True)  # warm-up
times = []
for _ in range(3):
    t0 = time.perf_counter()
    scores, status, bounds, ops, ops_off = ctx.bio_align_batch(sc, arena, tab, True)
    times.append(time.perf_counter() - t0)
assert (status == 0).all()
dt = min(times)
cells = float((tab["pattern_len"].astype(np.int64) * tab["text_len"].astype(np.int64)).sum())
osc = O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0)
done, t_cpu, ccells = 0, 0.0, 0.0
while done < n_pairs and t_cpu < cpu_s:
    x, y = cases[done]
    t0 = time.perf_counter()
    s = O.bio_custom_align(osc, x, y)[0]
    t_cpu += time.perf_counter() - t0
    assert s == scores[done], (done, s, scores[done])
    ccells += len(x) * len(y)
    done += 1
# ---- K5: the score-only kernel (what merge_pools.rs consumes) ----
ctx.bio_score_batch(sc, arena, tab[:8])  # warm-up
t5, k5_kernel_ms = [], []
for _ in range(5):
    t0 = time.perf_counter()
    s5, st5 = ctx.bio_score_batch(sc, arena, tab)
    t5.append(time.perf_counter() - t0)
    k5_kernel_ms.append(ctx.stats()["kernel_ms"])
assert (st5 == 0).all() and (s5 == scores).all(), "K5 scores differ from K4"
dt5, km5 = min(t5), min(k5_kernel_ms)
int_peak = ctx.measure_int_peak()[0]
k5 = {"value": n_pairs / dt5, "unit": "alignments/s", "ms": dt5 * 1e3, "gcups": cells / dt5 / 1e9, "kernel_ms": km5,
      "kernel_gcups": cells / km5 / 1e6, "includes": "H2D + kernel + D2H through sfc_bio_score_batch",
      "roofline_int": {"ops_per_cell": 8, "achieved_gops": 8 * cells / km5 / 1e6, "peak_gops": int_peak,
                       "frac": 8 * cells / km5 / 1e6 / int_peak}, "scores_identical_to_k4": True}
norm_ok = int(((scores / tab["pattern_len"]) >= 0.97).sum())
print(json.dumps({"metric": "merge_alignments_per_s", "value": n_pairs / dt, "unit": "alignments/s", "pairs": n_pairs,
                  "ms": dt * 1e3, "gcups": cells / dt / 1e9, "includes": "H2D + kernel + D2H through sfc_bio_align_batch",
                  "cpu_baseline": {"value": done / t_cpu, "unit": "alignments/s", "cores": 1, "kind": "port", "gcups": ccells / t_cpu / 1e9,
                                   "sample": f"first {done} pairs, scores identical"},
                  "pairs_with_norm_score_ge_0.97": norm_ok, "k5_score_only": k5}))
