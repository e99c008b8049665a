#!/usr/bin/env python
"""bench.py — read x haplotype alignments/s (and GCUPS-equivalent) of the WFA hot path on B200.

One step = one pass of the hot path over one batch of synthetic pairs.  Default workload = BASELINE.json
configs[1]: HiFi read segments 2-20 kb vs 2 allele haplotypes each, 0.1 % error, 35 bp-5 kb indel SVs, aligned
with sawfish's gap-affine-2-piece preset (new_large_indel_aligner) WITH CIGAR, sequences reversed, clip 0.4.

  value  = whole-job alignments/s with the batch already resident in HBM (kernels only, CUDA events on the
           library's stream, max over ranks);
  e2e    = the same through the C ABI from pinned HOST buffers (H2D + pack + kernels + D2H of scores/ops);
  roofline / roofline_int / cpu_baseline: see DESIGN.md §Measurement.

`--impl reference` times the CPU restatement of the reference path (oracle/, all host threads) instead.
Multi-GPU (torchrun, one rank per GPU): pairs are sharded by batch, no data-path collective ("weak" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (generator kwargs, preset, want_cigar, description)
    "config2": dict(preset="large_indel", want_cigar=True,
                    desc="configs[1]: read segments 2-20 kb x 2 allele haplotypes, 0.1% err, 35bp-5kb indel SV, gap-affine-2p + CIGAR"),
    "config2_linear": dict(preset="consensus", want_cigar=False,
                           desc="configs[1] inputs with the gap-linear consensus preset, score only"),
    "config2b": dict(preset="consensus", want_cigar=False,
                     desc="score_sv-shaped: 100-500 bp read flank x ref/alt allele flanks, gap-linear, score only"),
    "config3": dict(preset="large_indel", want_cigar=True,
                    desc="refine_sv-shaped: contig vs reference window (+2-region N spacer), gap-affine-2p + CIGAR"),
    "config4": dict(preset="consensus", want_cigar=False,
                    desc="configs[3] in miniature: 8-sample joint-call, per SV group x sample x read x breakend x registration x allele "
                         "flank pairs (100-500 bp, truncations, overlapping haplotypes), gap-linear, score only; --reads = SV groups per GPU"),
    "config5": dict(preset="large_indel", want_cigar=True,
                    desc="stress: 10-50 kb insertion/duplication haplotypes at 1-5% divergence"),
}
DEFAULT_READS = {"config4": 96, "config2": 1024, "config2_linear": 1024, "config2b": 200000, "config3": 512, "config5": 8}


def make_batch(workload, n_reads, seed):
    from sawfish_b200 import synth
    if workload in ("config2", "config2_linear"):
        arena, pairs, _ = synth.config2_read_vs_haplotypes(n_reads, seed=seed)
    elif workload == "config2b":
        arena, pairs, _ = synth.config2b_score_sv_flanks(n_reads, seed=seed)
    elif workload == "config3":
        arena, pairs = synth.config3_refine_contigs(n_reads, seed=seed)
    elif workload == "config4":
        from sawfish_b200 import score_sv
        builder, _meta = score_sv.synth_sv_groups(n_reads, n_samples=8, depth=30, seed=seed)
        arena, pairs = builder.finish()
    else:
        arena, pairs = synth.config5_stress(n_reads, seed=seed)
    return arena, pairs


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, n_pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed ncu
    --set full capture of this exact command (profiles/ncu_traffic.json); None when no capture matches."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    for rec in json.load(open(p)):
        if rec["workload"] == workload and rec["pairs_per_gpu"] == n_pairs:
            return rec["dram_bytes_per_launch"]
    return None


def algorithmic_bytes(pairs, n_ops_words):
    """Bytes the path must move per batch: packed inputs (4 bit/base), 8 B score+status, run-length ops."""
    bases = int(pairs["pattern_len"].astype(np.int64).sum() + pairs["text_len"].astype(np.int64).sum())
    return bases // 2 + 8 * len(pairs) + 4 * int(n_ops_words)


def cpu_reference_run(workload, arena, pairs, budget_s, threads=0):
    """Times the CPU restatement (oracle) on a bounded sample of the batch: consecutive slices of the batch
    (in batch order, so the sample has the batch's own mix of matching / non-matching haplotypes) are aligned
    until ~budget_s seconds of wall time have been spent."""
    from oracle import oracle as O
    cfg = WORKLOADS[workload]
    pen = O.CONSENSUS if cfg["preset"] == "consensus" else O.LARGE_INDEL
    if not threads:
        # all host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every
        # rank, which would turn the N>1 reference arm into a one-thread baseline.
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or O.max_threads()
    opairs = pairs.astype(O.PAIR_DTYPE)
    n = len(opairs)
    done, secs, cells = 0, 0.0, 0
    chunk = max(2 * threads, 8)
    scores = np.zeros(n, np.int32)
    while done < n and secs < budget_s:
        m = min(n - done, chunk)
        t0 = time.perf_counter()
        sc, status, stats, _, _ = O.align_batch(pen, arena, opairs[done:done + m], want_ops=True, mode=O.MODE_COMPACT, n_threads=threads)
        dt = time.perf_counter() - t0
        assert (status == 0).all()
        scores[done:done + m] = sc
        cells += int(stats["cells"].sum())
        done += m
        secs += dt
        if dt < budget_s / 8:
            chunk *= 2
    dp = float((opairs["pattern_len"][:done].astype(np.int64) * opairs["text_len"][:done].astype(np.int64)).sum())
    return {"pairs": done, "seconds": secs, "value": done / secs, "cores": threads, "cells": cells,
            "gcups": dp / secs / 1e9, "scores": scores[:done]}


def run_reference(args, rank, world):
    if rank != 0:
        return
    arena, pairs = make_batch(args.workload, args.reads, args.seed)
    per_step_budget = max(2.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    info = None
    for it in range(args.warmup + args.steps):
        info = cpu_reference_run(args.workload, arena, pairs, per_step_budget)
        if it >= args.warmup:
            vals.append(info)
    tot_pairs = sum(v["pairs"] for v in vals)
    tot_s = sum(v["seconds"] for v in vals)
    val = tot_pairs / tot_s
    sample = f"first {info['pairs']} of {len(pairs)} pairs of the batch per step (~{per_step_budget:.0f} s budget)"
    out = {"impl": "reference", "metric": "alignments_per_s", "value": val, "unit": "alignments/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": {"workload": WORKLOADS[args.workload]["desc"], "name": args.workload, "reads_per_gpu": args.reads,
                      "pairs_per_gpu": int(len(pairs))},
           "gcups": sum(v["gcups"] * v["seconds"] for v in vals) / tot_s,
           "cpu_baseline": {"value": val, "unit": "alignments/s", "cores": info["cores"], "kind": "port", "sample": sample,
                            "note": "WFA2-semantics CPU restatement (oracle/), not the sawfish binary: no Rust toolchain / WFA2-lib here"},
           "e2e": {"value": val, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_REAL_STDOUT = None


def claim_stdout():
    """Rank 0 must print exactly ONE JSON line on stdout.  NCCL prints its version banner there (whenever NCCL_DEBUG is
    VERSION or higher) and other libraries may chat too, so file descriptor 1 is pointed at stderr for the whole run and
    the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="reads (or clusters/pairs) per GPU per step")
    ap.add_argument("--seed", type=int, default=0x5AF15 + 2)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU baseline work (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="check GPU scores of the CPU-baseline sample against the oracle")
    args = ap.parse_args()
    if args.reads <= 0:
        args.reads = DEFAULT_READS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from sawfish_b200.batch import AlignContext, CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES
    from sawfish_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints exactly one JSON line on stdout: NCCL's version banner / debug output (printed to stdout
        # whenever NCCL_DEBUG is VERSION or higher, WARN included) goes to a file instead
        os.environ["NCCL_DEBUG_FILE"] = "/tmp/sfc_bench_nccl.%h.%p.log"
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"  # the banner is printed at VERSION and WARN; INFO / TRACE are left to the user
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = WORKLOADS[args.workload]
    pen = CONSENSUS_PENALTIES if cfg["preset"] == "consensus" else LARGE_INDEL_PENALTIES
    want_cigar = cfg["want_cigar"]
    # every rank aligns its own shard of SV-candidate batches.  Weak scaling: per-GPU work is held FIXED by giving
    # every rank an identically generated shard (same generator seed); a different seed per rank would change the
    # heaviest alignment of the shard and with it the per-rank time.
    arena, pairs = make_batch(args.workload, args.reads, args.seed)
    n_pairs = len(pairs)
    ctx = AlignContext(local_rank)

    # pinned host staging for the end-to-end leg
    arena_pin = torch.from_numpy(arena).pin_memory()
    pairs_pin = torch.from_numpy(pairs.view(np.uint8)).pin_memory()
    scores_pin = torch.empty(n_pairs, dtype=torch.int32).pin_memory()
    status_pin = torch.empty(n_pairs, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def e2e_step():
        ctx.upload_ptr(arena_pin.data_ptr(), arena_pin.numel(), pairs_pin.data_ptr(), n_pairs)
        ctx.run(pen, want_cigar)
        return ctx.download(out=(scores_pin.numpy(), status_pin.numpy()))

    # ---- warm-up (also allocates device workspace) ----
    # At least W (>= 3) steps AND at least ~2.5 s: the first GPU process on a fresh box runs its first few hundred
    # milliseconds measurably slower (140 vs 102 ms per config2 step was seen with a 0.4 s warm-up), and the driver's
    # bench is exactly that process.  The count actually done is what the JSON line reports as "warmup".
    n_warm = 0
    t_warm = time.perf_counter()
    while n_warm < max(3, args.warmup) or (time.perf_counter() - t_warm < 2.5 and n_warm < 64):
        scores, status, ops, ops_off = e2e_step()
        n_warm += 1
    assert (status == 0).all(), f"pair status != 0: {np.unique(status)}"
    n_ops_words = 0 if ops is None else len(ops)
    st0 = ctx.stats()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident throughput: kernels only ----
    ctx.upload_ptr(arena_pin.data_ptr(), arena_pin.numel(), pairs_pin.data_ptr(), n_pairs)
    barrier()
    kern_ms = 0.0
    launches = 0
    cells = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations
        torch.cuda.synchronize()
        ctx.run(pen, want_cigar)
        torch.cuda.synchronize()
        # kernel_ms is measured by CUDA events on the library's own stream around the launches of this run
        ctx.download(out=(scores_pin.numpy(), status_pin.numpy()))
        s = ctx.stats()
        kern_ms += s["kernel_ms"]
        launches += s["launches"] - 1  # minus the pack kernel, which belongs to upload
        cells = s["cells"]
    barrier()
    wall_resident = time.perf_counter() - t0
    # ---- end to end through the C ABI from pinned host buffers ----
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2e_dev_ms = 0.0
    for _ in range(args.steps):
        e2e_step()
        s = ctx.stats()
        e2e_dev_ms += s["h2d_ms"] + s["pack_ms"] + s["kernel_ms"] + s["d2h_ms"]
    barrier()
    e2e_wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st = ctx.stats()

    # max over ranks (device time), whole-job aggregate
    t = torch.tensor([kern_ms, e2e_wall * 1e3, e2e_dev_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(n_pairs), float(synth.gcups_cells(pairs)), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    kern_ms_max, e2e_wall_ms_max, e2e_dev_ms_max = [float(x) for x in t.tolist()]
    all_pairs, all_cells_dp, all_launches = [float(x) for x in tot.tolist()]
    ms_per_step = kern_ms_max / args.steps
    value = all_pairs / (ms_per_step * 1e-3)
    e2e_value = all_pairs / (e2e_wall_ms_max * 1e-3 / args.steps)

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg_bytes = algorithmic_bytes(pairs, n_ops_words)
        achieved = alg_bytes / (kern_ms / args.steps * 1e-3) / 1e9
        int_peak_gops, _ = ctx.measure_int_peak()
        ops_per_cell = 26 if cfg["preset"] == "large_indel" else 12
        int_ops = ops_per_cell * cells
        out = {
            "metric": "alignments_per_s", "value": value, "unit": "alignments/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "name": args.workload, "reads_per_gpu": args.reads, "pairs_per_gpu": n_pairs,
                       "preset": cfg["preset"], "want_cigar": want_cigar, "l2": "256 MiB flush between timed steps",
                       "sharding": "independent pair batches per rank, host-side gather, no collective"},
            "gcups": all_cells_dp / (ms_per_step * 1e-3) / 1e9,
            "wavefront_cells_per_step": int(cells),
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "alignments/s", "h2d_bytes_per_step": int(st["h2d_bytes"]),
                    "d2h_bytes_per_step": int(st["d2h_bytes"]), "ms_per_step_wall": e2e_wall_ms_max / args.steps,
                    "ms_per_step_device": e2e_dev_ms_max / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload, n_pairs), "peak_source": peak_src,
                         "note": "algorithmic bytes = 4-bit packed inputs + scores/status + run-length ops; the path is integer-ALU/latency bound, see roofline_int"},
            "roofline_int": {"bound": "int32-alu", "achieved": int_ops / (kern_ms / args.steps * 1e-3) / 1e9, "peak": int_peak_gops,
                             "unit": "Gop/s", "frac": int_ops / (kern_ms / args.steps * 1e-3) / 1e9 / int_peak_gops,
                             "ops_per_cell": ops_per_cell, "peak_source": "IADD/LOP3 micro-benchmark run in this process (sfc_measure_int_peak)"},
            "kernel_split": {"n_score_kernel_pairs": st["n_score_kernel"], "n_align_kernel_pairs": st["n_align_kernel"],
                             "retries": st["retries"], "pack_ms": st["pack_ms"], "h2d_ms": st["h2d_ms"], "d2h_ms": st["d2h_ms"]},
            "wall_ms_per_step_resident": wall_resident * 1e3 / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(args.workload, arena, pairs, args.cpu_budget)
            if args.verify:
                assert (cb["scores"] == scores[:cb["pairs"]]).all(), "GPU scores differ from the oracle on the CPU sample"
            out["cpu_baseline"] = {"value": cb["value"], "unit": "alignments/s", "cores": cb["cores"], "kind": "port",
                                   "sample": f"first {cb['pairs']} of {n_pairs} pairs of the step batch, {cb['seconds']:.1f} s",
                                   "gcups": cb["gcups"],
                                   "note": "WFA2-semantics CPU restatement (oracle/, OpenMP over pairs, fresh aligner state per pair)"}
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
