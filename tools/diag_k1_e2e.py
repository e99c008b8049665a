"""Diagnostic: host-side time of upload / run / download for the score_sv-shaped batch (config2b) next to the device times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sawfish_b200.batch import AlignContext, CONSENSUS_PENALTIES as PEN
ctx = AlignContext(0)
arena, pairs, _ = bench.make_batch("config2b", 200000, 0x5AF15 + 2)
a = torch.from_numpy(arena).pin_memory(); p = torch.from_numpy(pairs.view(np.uint8).copy()).pin_memory(); n = len(pairs)
sc = torch.empty(n, dtype=torch.int32).pin_memory(); st = torch.empty(n, dtype=torch.int32).pin_memory()
rows = []
for it in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ctx.upload_ptr(a.data_ptr(), a.numel(), p.data_ptr(), n)
    t1 = time.perf_counter(); ctx.run(PEN, False)
    t2 = time.perf_counter(); ctx.download(out=(sc.numpy(), st.numpy()))
    t3 = time.perf_counter(); s = ctx.stats()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, s["h2d_ms"], s["pack_ms"], s["kernel_ms"], s["d2h_ms"]))
print(n, "pairs; per step: upload() run() download() total | h2d pack kernel d2h (ms)")
for r in rows[2:]:
    print(" ".join(f"{v:7.2f}" for v in r))
