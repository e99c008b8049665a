#!/usr/bin/env python
"""Time every pair of a synthetic workload alone on one CTA (SFC_K2_CLUSTER=1) to find outliers."""
import os, sys, time
import numpy as np
os.environ.setdefault("SFC_K2_CLUSTER", "1")
from sawfish_b200 import synth, batch
arena, pairs = synth.config3_refine_contigs(int(sys.argv[1]) if len(sys.argv) > 1 else 512)
ctx = batch.AlignContext(0)
pen = batch.LARGE_INDEL_PENALTIES
ctx.align_batch(pen, arena, pairs[:4], True)
ts = []
for i in range(len(pairs)):
    t = time.perf_counter()
    sc, st, ops, off = ctx.align_batch(pen, arena, pairs[i:i + 1], True)
    ts.append(time.perf_counter() - t)
ts = np.array(ts) * 1e3
o = np.argsort(-ts)
print("sum ms %.1f  avg %.2f  max %.2f" % (ts.sum(), ts.mean(), ts.max()))
for i in o[:12]:
    print(i, "plen", pairs["pattern_len"][i], "tlen", pairs["text_len"][i], "ms %.2f" % ts[i])
print("quantiles", np.quantile(ts, [.5, .9, .99]).round(2))
