"""Diagnostic: per-step kernel_ms of the config2 batch in different call patterns (why e2e kernel time > resident)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sawfish_b200.batch import AlignContext, LARGE_INDEL_PENALTIES as PEN
ctx = AlignContext(0)
batches = [bench.make_batch("config2", 1024, 0x5AF15 + 2 + 1000 * b) for b in range(3)]
pins = []
for arena, pairs, _ in batches:
    pins.append((torch.from_numpy(arena).pin_memory(), torch.from_numpy(pairs.view(np.uint8).copy()).pin_memory(), len(pairs)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def up(b):
    a, p, n = pins[b]; ctx.upload_ptr(a.data_ptr(), a.numel(), p.data_ptr(), n)
def step(b, mode):
    up(b)
    if mode == "flush": flush.fill_(1); torch.cuda.synchronize()
    elif mode == "sync": torch.cuda.synchronize()
    t0 = time.perf_counter(); ctx.run(PEN, True); t1 = time.perf_counter(); r = ctx.download(); t2 = time.perf_counter()
    s = ctx.stats()
    return s["kernel_ms"], (t1 - t0) * 1e3, (t2 - t1) * 1e3, s["launches"], s["retries"]
for _ in range(6): step(_ % 3, "none")
for mode in ("none", "sync", "flush", "none", "flush"):
    out = [step(i % 3, mode) for i in range(6)]
    print(mode, " kernel_ms:", [round(o[0], 1) for o in out], " run() host ms:", [round(o[1], 1) for o in out], " download ms:", [round(o[2], 1) for o in out], out[0][3:])
