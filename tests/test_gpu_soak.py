"""The parity soak as a tracked GPU test (60 s budget): seeded batches of every shape through the C ABI against the
oracle, bit-exact scores and run-length ops, both ring widths (int32 / int16).  compute-sanitizer is closed on this pool,
so this is also the memory-safety net of K1's unguarded null-cell loads and K2's plan / termination protocol."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_soak_60s(ctx):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    import soak
    n_batches, n_pairs, last_seed = soak.run(60.0, ctx)
    print(f"soak: {n_batches} batches, {n_pairs} pairs, seeds 1001..{last_seed}")
    assert n_batches >= 12 and n_pairs > 0
