#!/usr/bin/env python
"""Run bench.py with the given args and print a one-line summary (or the error tail)."""
import json
import subprocess
import sys

p = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
try:
    d = json.loads(line)
    cb = d.get("cpu_baseline", {})
    print(d["config"]["name"], "pairs", d["config"].get("pairs_per_gpu"), "n_gpus", d["n_gpus"], "| value", round(d["value"], 1),
          "ms", round(d["ms_per_step"], 2), "| e2e", round(d["e2e"]["value"], 1), "| gcups", round(d.get("gcups", 0), 1),
          "| cells", d.get("wavefront_cells_per_step"), "| int_frac", round(d.get("roofline_int", {}).get("frac", 0), 4),
          "| launches", d.get("gpu_launches"), "| cpu", round(cb.get("value", 0), 1), cb.get("cores"), "| retries",
          d.get("kernel_split", {}).get("retries"), "| clk", d.get("clocks", {}).get("sm_mhz"))
except Exception as e:
    print("BENCH FAILED rc", p.returncode, e)
    print(p.stdout[-1500:])
    print(p.stderr[-3000:])
