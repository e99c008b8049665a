#!/usr/bin/env python
"""End-to-end throughput of ONE GPU fed by several host threads, each with its own sfc_ctx (the library's threading model:
one context per (host thread, GPU), no global state — what sawfish's rayon workers would hold).

A single context runs upload -> kernels -> download back to back, so its end-to-end rate is below the kernel rate by the
H2D / pack / host-planning time.  With two or three contexts on the same GPU those phases of one batch overlap the kernels
of another; this tool measures how much of the gap that recovers.  Every step is a full pass from pinned HOST buffers
(H2D + pack + kernels + D2H).  usage: python tools/bench_pipeline.py [workload] [steps_per_thread]"""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sawfish_b200.batch import AlignContext, CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "config2b"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = bench.WORKLOADS[workload]
pen = CONSENSUS_PENALTIES if cfg["preset"] == "consensus" else LARGE_INDEL_PENALTIES
arena, pairs, _groups = bench.make_batch(workload, bench.DEFAULT_READS[workload], 0x5AF15 + 2)
n = len(pairs)
arena_pin = torch.from_numpy(arena).pin_memory()
pairs_pin = torch.from_numpy(pairs.view(np.uint8)).pin_memory()
out = {"metric": "alignments_per_s", "unit": "alignments/s", "workload": workload, "pairs_per_step": int(n),
       "steps_per_thread": steps, "includes": "H2D + pack + kernels + D2H from pinned host buffers, every step", "runs": []}
ref_scores = None
for n_ctx in (1, 2, 3):
    ctxs = [AlignContext(0) for _ in range(n_ctx)]
    bufs = [(torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()) for _ in range(n_ctx)]

    def step(i):
        ctxs[i].upload_ptr(arena_pin.data_ptr(), arena_pin.numel(), pairs_pin.data_ptr(), n)
        ctxs[i].run(pen, cfg["want_cigar"])
        return ctxs[i].download(out=(bufs[i][0].numpy(), bufs[i][1].numpy()))

    for i in range(n_ctx):  # warm-up: workspaces
        for _ in range(3):
            step(i)
    torch.cuda.synchronize()
    barrier = threading.Barrier(n_ctx + 1)

    def worker(i):
        barrier.wait()
        for _ in range(steps):
            step(i)
        barrier.wait()

    th = [threading.Thread(target=worker, args=(i,)) for i in range(n_ctx)]
    for t in th:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    sc = bufs[0][0].numpy().copy()
    assert (bufs[0][1].numpy() == 0).all()
    if ref_scores is None:
        ref_scores = sc
    assert (sc == ref_scores).all()
    out["runs"].append({"contexts": n_ctx, "value": n_ctx * steps * n / dt, "ms_per_step_effective": dt * 1e3 / (n_ctx * steps)})
    for c in ctxs:
        c.close()
for r in out["runs"]:
    r["vs_one_context"] = r["value"] / out["runs"][0]["value"]
print(json.dumps(out))
