"""Diagnostic: step-to-step variance of the config2 step next to 20 ms nvidia-smi samples (clocks, power, throttle reasons)."""
import sys, os, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sawfish_b200.batch import AlignContext, LARGE_INDEL_PENALTIES as PEN
q = "timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,utilization.gpu"
POLL = os.environ.get("DIAG_POLL_MS", "20")
proc = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", POLL], stdout=subprocess.PIPE, text=True) if POLL != "0" else subprocess.Popen(["sleep", "1000"], stdout=subprocess.PIPE, text=True)
lines = []
def rd():
    for ln in proc.stdout:
        lines.append((time.time(), ln.strip()))
threading.Thread(target=rd, daemon=True).start()
ctx = AlignContext(0)
batches = [bench.make_batch("config2", 1024, 0x5AF15 + 2 + 1000 * b) for b in range(3)]
pins = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(p.view(np.uint8).copy()).pin_memory(), len(p)) for a, p, _ in batches]
steps = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 50):
    a, p, n = pins[it % 3]
    ctx.upload_ptr(a.data_ptr(), a.numel(), p.data_ptr(), n)
    t0 = time.time(); ctx.run(PEN, True); t1 = time.time(); ctx.download()
    steps.append((t0, t1, ctx.stats()["kernel_ms"], (t1 - t0) * 1e3))
proc.terminate()
for t0, t1, km, hm in steps[6:]:
    smp = [ln for t, ln in lines if t0 - 0.02 <= t <= t1 + 0.02]
    flag = "SLOW" if km > 95 else ""
    print(f"{km:7.1f} ms (host {hm:6.1f}) {flag:4s} | " + " ; ".join(",".join(s.split(",")[1:]) for s in smp[:6]))
