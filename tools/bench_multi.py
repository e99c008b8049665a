#!/usr/bin/env python
"""One process, several GPUs: strong scaling of ONE joint-call-shaped batch through sfc_multi_align_batch.

BASELINE config 4 in miniature (8 samples, SV groups x reads x breakends x registrations x alleles, gap-linear, score only).
The batch is sharded by SV group (the reference's rayon task, src/joint_call/joint_call_all_samples.rs:127-142) with
sfc_shard_plan over 1, 2, 4, ... of the visible GPUs, each shard runs on its GPU's submitter thread, results are gathered
on the host in the caller's pair order; no collective.  Timed end to end from HOST buffers (host sharding + H2D + kernels
+ D2H + gather), best of `reps`.  Scores must be identical for every GPU count.
usage: python tools/bench_multi.py [n_groups] [reps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sawfish_b200 import _lib, score_sv  # noqa: E402
from sawfish_b200.batch import CONSENSUS_PENALTIES, MultiGpuContext  # noqa: E402

n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_samples = 8
t0 = time.perf_counter()
builder, meta = score_sv.synth_sv_groups(n_groups, n_samples=n_samples, depth=30, seed=0x5AF15 + 4)
arena, pairs = builder.finish()
groups = np.array([tag // n_samples for tag, _slot in builder.slots], np.uint32)
gen_s = time.perf_counter() - t0
n_dev = _lib.lib().sfc_device_count()
out = {"metric": "alignments_per_s", "unit": "alignments/s", "scaling": "strong", "pairs": int(len(pairs)), "sv_groups": n_groups,
       "samples": n_samples, "workload": "config4-shaped (score_sv flank pairs), consensus preset, score only", "gen_s": gen_s,
       "includes": "host sharding by SV group + H2D + kernels + D2H + host gather, one process", "runs": []}
ref = None
for n in [g for g in (1, 2, 4, 8) if g <= n_dev]:
    multi = MultiGpuContext(list(range(n)))
    multi.align_batch(CONSENSUS_PENALTIES, arena, pairs[:4096], False, group_of_pair=groups[:4096])  # warm-up: contexts, workspaces
    scores, status, _, _ = multi.align_batch(CONSENSUS_PENALTIES, arena, pairs, False, group_of_pair=groups)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        scores, status, _, _ = multi.align_batch(CONSENSUS_PENALTIES, arena, pairs, False, group_of_pair=groups)
        times.append(time.perf_counter() - t0)
    assert (status == 0).all()
    if ref is None:
        ref = scores.copy()
    assert (scores == ref).all(), f"scores differ between 1 and {n} GPUs"
    dt = min(times)
    km = multi.last_kernel_ms()
    out["runs"].append({"n_gpus": n, "value": len(pairs) / dt, "ms": dt * 1e3, "kernel_ms_per_gpu": [round(float(x), 2) for x in km],
                        "speedup_vs_1": None})
    multi.close()
for r in out["runs"]:
    r["speedup_vs_1"] = r["value"] / out["runs"][0]["value"]
print(json.dumps(out))
