#!/usr/bin/env python
"""Phase timers (SFC_K2_PROF) of single pairs of config3, each alone on one CTA."""
import os, sys
import numpy as np
os.environ.setdefault("SFC_K2_CLUSTER", "1")
os.environ["SFC_K2_PROF"] = "1"
from sawfish_b200 import synth, batch
arena, pairs = synth.config3_refine_contigs(512)
ctx = batch.AlignContext(0)
pen = batch.LARGE_INDEL_PENALTIES
for i in [int(x) for x in sys.argv[1:]]:
    sys.stderr.write("pair %d plen %d tlen %d\n" % (i, pairs["pattern_len"][i], pairs["text_len"][i]))
    sc, st, ops, off = ctx.align_batch(pen, arena, pairs[i:i + 1], True)
    sys.stderr.write("  score %d status %d stats %s\n" % (sc[0], st[0], ctx.stats()))
