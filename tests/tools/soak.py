#!/usr/bin/env python
"""Soak test: many seeded batches of every shape through the CUDA path vs the oracle (run on the GPU box).
usage: python tests/tools/soak.py [seconds]
`run(budget, ctx)` is also what tests/test_gpu_soak.py calls with a 60 s budget (pytest -m gpu)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import compare_with_oracle  # noqa: E402


def run(budget, ctx, seed0=1000, widths=(False, True)):
    """Seeded batches of every shape (random small pairs in LITERAL mode, low complexity, score_sv flanks, refine contigs,
    long reads, stress) until `budget` seconds have passed; every batch bit-exact (scores and run-length ops) against the
    oracle, in both ring widths.  Returns (batches, pairs, last seed)."""
    from oracle import oracle as O
    from sawfish_b200 import synth
    from sawfish_b200.batch import CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES
    t0 = time.time()
    seed = seed0
    n_batches = n_pairs = 0
    had = os.environ.get("SFC_K2_NARROW")
    try:
        while time.time() - t0 < budget:
            seed += 1
            kind = seed % 6
            if kind == 0:
                arena, pairs = synth.random_pairs(800, seed=seed, max_len=200)
                jobs = [(CONSENSUS_PENALTIES, False), (CONSENSUS_PENALTIES, True), (LARGE_INDEL_PENALTIES, True)]
                mode = O.MODE_LITERAL
            elif kind == 1:
                arena, pairs = synth.low_complexity_pairs(60, seed=seed, len_lo=200, len_hi=2500)
                jobs = [(CONSENSUS_PENALTIES, False), (LARGE_INDEL_PENALTIES, True)]
                mode = None
            elif kind == 2:
                arena, pairs, _ = synth.config2b_score_sv_flanks(1500, seed=seed)
                jobs = [(CONSENSUS_PENALTIES, False)]
                mode = None
            elif kind == 3:
                arena, pairs = synth.config3_refine_contigs(int(np.random.default_rng(seed).integers(1, 40)), seed=seed, two_region_frac=0.4)
                jobs = [(LARGE_INDEL_PENALTIES, True)]
                mode = None
            elif kind == 4:
                arena, pairs, _ = synth.config2_read_vs_haplotypes(int(np.random.default_rng(seed).integers(1, 6)), seed=seed, len_lo=1000,
                                                                    len_hi=9000, sv_hi=2500)
                jobs = [(LARGE_INDEL_PENALTIES, True), (LARGE_INDEL_PENALTIES, False)]
                mode = None
            else:
                arena, pairs = synth.config5_stress(1, seed=seed, len_lo=1500, len_hi=4000, div_lo=0.01, div_hi=0.05)
                jobs = [(LARGE_INDEL_PENALTIES, True), (CONSENSUS_PENALTIES, True)]
                mode = None
            for narrow in widths:
                if narrow:
                    os.environ["SFC_K2_NARROW"] = "1"
                else:
                    os.environ.pop("SFC_K2_NARROW", None)
                for pen, wc in jobs:
                    compare_with_oracle(O, ctx, pen, arena, pairs, want_cigar=wc, mode=mode)
                    n_batches += 1
                    n_pairs += len(pairs)
    finally:
        if had is None:
            os.environ.pop("SFC_K2_NARROW", None)
        else:
            os.environ["SFC_K2_NARROW"] = had
    return n_batches, n_pairs, seed


if __name__ == "__main__":
    from sawfish_b200.batch import AlignContext
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    t0 = time.time()
    ctx = AlignContext(0)
    nb, npairs, seed = run(budget, ctx)
    print(f"soak ok: {nb} batches, {npairs} pairs, seeds 1001..{seed}, {time.time() - t0:.0f} s")
