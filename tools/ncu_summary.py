#!/usr/bin/env python
"""Print the key metrics of an .ncu-rep (first launch) and an opcode histogram of the hot SASS."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
for h, u, v in zip(hdr, units, vals):
    if any(h == w or (w == 'launch__cluster' and h.startswith(w)) for w in want):
        print(f"- {h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data)
tots = sum(int(r[isamp]) for r in data)
op, ops = collections.Counter(), collections.Counter()
for r in data:
    srcl = r[isrc].strip()
    if not srcl:
        continue
    t = srcl.split()
    o = (t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]
    op[o] += int(r[ia]); ops[o] += int(r[isamp])
print(f"- total warp instructions {tot}, stall samples {tots}")
print("- opcode mix (executed % / stall-sample %): " + ", ".join(f"{o} {c / tot * 100:.1f}/{ops[o] / max(1, tots) * 100:.1f}" for o, c in op.most_common(14)))
