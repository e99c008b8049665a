#!/usr/bin/env python
"""bench.py — read x haplotype alignments/s (and GCUPS-equivalent) of the WFA hot path on B200.

One step = one pass of the hot path over one batch of synthetic pairs.  Default workload = BASELINE.json
configs[1]: HiFi read segments 2-20 kb vs 2 allele haplotypes each, 0.1 % error, 35 bp-5 kb indel SVs, aligned
with sawfish's gap-affine-2-piece preset (new_large_indel_aligner) WITH CIGAR, sequences reversed, clip 0.4.

  value  = whole-job alignments/s with the batch already resident in HBM (kernels only, CUDA events on the
           library's stream, max over ranks);
  e2e    = the same through the C ABI from pinned HOST buffers (H2D + pack + kernels + D2H of scores/ops; at N > 1
           also the shard plan, the compaction of the rank's shard and the host-side gather on rank 0);
  roofline / roofline_int / cpu_baseline / verified / secondary: see DESIGN.md §Measurement.

Every timed step aligns a DIFFERENT batch (seed + 1000 * (step mod n_batches)); the first batch is also the one the
CPU restatement is timed on, and the GPU's scores AND run-length ops of that sample are compared with it (`verified`).

Multi-GPU (torchrun, one rank per GPU): every rank builds the same GLOBAL batch of N x reads-per-GPU SV groups, cuts it
with the library's LPT sharder (sfc_estimate_pair_work + sfc_shard_plan, sharding.shard_pairs), aligns only its own
shard, and rank 0 gathers the results on the host in original pair order.  No data-path collective.  Per-GPU work is
fixed as N grows ("weak"); value = global pairs / max-over-ranks device time.

`--impl reference` times the CPU restatement of the reference path (oracle/, all host threads) instead.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: preset, want_cigar, description, dominant kernel source (for the ncu traffic stamp)
    "config2": dict(preset="large_indel", want_cigar=True, kernel="sfc_k_align2.cuh",
                    desc="configs[1]: read segments 2-20 kb x 2 allele haplotypes, 0.1% err, 35bp-5kb indel SV, gap-affine-2p + CIGAR"),
    "config2_linear": dict(preset="consensus", want_cigar=False, kernel="sfc_k_align.cuh",
                           desc="configs[1] inputs with the gap-linear consensus preset, score only"),
    "config2b": dict(preset="consensus", want_cigar=False, kernel="sfc_k_score.cuh",
                     desc="score_sv-shaped: 100-500 bp read flank x ref/alt allele flanks, gap-linear, score only"),
    "config3": dict(preset="large_indel", want_cigar=True, kernel="sfc_k_align2.cuh",
                    desc="refine_sv-shaped: contig vs reference window (+2-region N spacer), gap-affine-2p + CIGAR"),
    "config4": dict(preset="consensus", want_cigar=False, kernel="sfc_k_score.cuh",
                    desc="configs[3] in miniature: 8-sample joint-call, per SV group x sample x read x breakend x registration x allele "
                         "flank pairs (100-500 bp, truncations, overlapping haplotypes), gap-linear, score only; --reads = SV groups per GPU"),
    "config5": dict(preset="large_indel", want_cigar=True, kernel="sfc_k_align2.cuh",
                    desc="stress: 10-50 kb insertion/duplication haplotypes at 1-5% divergence"),
}
DEFAULT_READS = {"config4": 96, "config2": 1024, "config2_linear": 1024, "config2b": 200000, "config3": 2048, "config5": 8}
# secondary legs of the default line (N = 1): the score_sv shapes, i.e. K1 (smaller than their stand-alone defaults so
# that the default run stays within minutes)
SECONDARY = {"config2b": 100000, "config4": 64}
L2_NOTE = "256 MiB flush between timed steps"
SHARD_NOTE = ("global batch of n_gpus x reads_per_gpu SV groups cut by the LPT sharder (sfc_shard_plan) over estimated work; "
              "each rank aligns its shard, host-side gather on rank 0, no collective")


def make_batch(workload, n_reads, seed):
    """Returns (arena, pairs, group id per pair): the SV group / read / cluster is the unit of sharding."""
    from sawfish_b200 import synth
    if workload in ("config2", "config2_linear"):
        arena, pairs, _ = synth.config2_read_vs_haplotypes(n_reads, seed=seed)
        groups = np.arange(len(pairs)) // 2
    elif workload == "config2b":
        arena, pairs, _ = synth.config2b_score_sv_flanks(n_reads, seed=seed)
        groups = np.arange(len(pairs)) // 2
    elif workload == "config3":
        arena, pairs = synth.config3_refine_contigs(n_reads, seed=seed)
        groups = np.arange(len(pairs))
    elif workload == "config4":
        from sawfish_b200 import score_sv
        builder, meta = score_sv.synth_sv_groups(n_reads, n_samples=8, depth=30, seed=seed)
        arena, pairs = builder.finish()
        groups = np.array([tag for tag, _ in builder.slots], np.int64) // meta["n_samples"]
    else:
        arena, pairs = synth.config5_stress(n_reads, seed=seed)
        groups = np.arange(len(pairs))
    return arena, pairs, np.asarray(groups, np.int64)


_COPY_POOL = None


def _copy_pool():
    global _COPY_POOL
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=4)
    return _COPY_POOL


def compact_shard(arena, pairs, idx, out=None):
    """The sequences of pairs[idx] copied into an arena of their own (merged byte intervals, offsets remapped); straight
    into `out` (the pinned staging buffer) when given."""
    sub = pairs[idx].copy()
    starts = np.concatenate([sub["pattern_off"], sub["text_off"]]).astype(np.int64)
    ends = starts + np.concatenate([sub["pattern_len"], sub["text_len"]]).astype(np.int64)
    order = np.argsort(starts, kind="stable")
    s, e = starts[order], ends[order]
    run_end = np.maximum.accumulate(e)
    new_iv = np.ones(len(s), bool)
    new_iv[1:] = s[1:] > run_end[:-1]
    iv_start = s[new_iv]
    iv_end = np.maximum.reduceat(e, np.flatnonzero(new_iv))
    iv_len = iv_end - iv_start
    iv_new = np.concatenate([[0], np.cumsum(iv_len)[:-1]])
    # one slice copy per merged interval (an SV group's sequences are contiguous in the arena: ~ one interval per group)
    total = int(iv_len.sum()) if len(iv_len) else 0
    if total == 0:
        local = np.zeros(0, np.uint8)
    elif out is not None:
        # straight into the staging buffer, in <= 4 MB pieces on a few host threads (numpy copies release the GIL): the copy
        # is the whole cost of the compaction (tens of MB per shard)
        local = out[:total]
        jobs = []
        for a, b, d in zip(iv_start.tolist(), iv_end.tolist(), iv_new.tolist()):
            for o in range(0, b - a, 4 << 20):
                n = min(4 << 20, b - a - o)
                jobs.append((a + o, d + o, n))

        def copy(js):
            for j in js:
                local[j[1]:j[1] + j[2]] = arena[j[0]:j[0] + j[2]]
        # four bundles of about equal bytes (a task per interval would cost more in hand-offs than it copies)
        bundles, cur, acc = [], [], 0
        for j in jobs:
            cur.append(j)
            acc += j[2]
            if acc >= (total + 3) // 4:
                bundles.append(cur)
                cur, acc = [], 0
        if cur:
            bundles.append(cur)
        if len(bundles) > 1:
            list(_copy_pool().map(copy, bundles))
        else:
            copy(jobs)
    else:
        local = np.concatenate([arena[a:b] for a, b in zip(iv_start.tolist(), iv_end.tolist())])
    which_p = np.searchsorted(iv_start, sub["pattern_off"].astype(np.int64), side="right") - 1
    which_t = np.searchsorted(iv_start, sub["text_off"].astype(np.int64), side="right") - 1
    sub["pattern_off"] = (sub["pattern_off"].astype(np.int64) - iv_start[which_p] + iv_new[which_p]).astype(np.uint64)
    sub["text_off"] = (sub["text_off"].astype(np.int64) - iv_start[which_t] + iv_new[which_t]).astype(np.uint64)
    return local, sub


def merge_rank_results(parts, n_tot, want_cigar):
    """Host-side gather on rank 0: `parts` = one (pair indices, scores, run-length ops, ops offsets) per rank -> scores, ops and
    ops offsets in the global pair order.  Every pair must come from exactly one rank.  No per-pair Python loop: this runs
    inside the timed end-to-end step."""
    g_scores = np.zeros(n_tot, np.int32)
    seen = np.zeros(n_tot, bool)
    g_len = np.zeros(n_tot, np.int64)
    for idx, sc, op, off in parts:
        assert not seen[idx].any(), "a pair was assigned to two ranks"
        g_scores[idx] = sc
        seen[idx] = True
        if want_cigar:
            g_len[idx] = np.diff(np.asarray(off, np.int64))
    assert seen.all(), "a pair was assigned to no rank"
    if not want_cigar:
        return g_scores, None, None
    g_off = np.zeros(n_tot + 1, np.int64)
    np.cumsum(g_len, out=g_off[1:])
    g_ops = np.zeros(int(g_off[-1]), np.uint32)
    for idx, sc, op, off in parts:
        off = np.asarray(off, np.int64)
        lens = np.diff(off)
        tot = int(off[-1] - off[0])
        if tot:
            dest = np.repeat(g_off[:-1][idx] - off[:-1], lens) + np.arange(int(off[0]), int(off[0]) + tot)
            g_ops[dest] = np.asarray(op)[int(off[0]):int(off[0]) + tot]
    return g_scores, g_ops, g_off


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []   # (arrival time, csv line)
        self.t_mark = 0.0

    def mark(self):
        """The timed region starts now: only samples from here on are reported.  (The sampler itself is started BEFORE the
        warm-up: the first nvidia-smi on a fresh box takes seconds to initialise and slows the kernels running meanwhile —
        measured: 102 ms per config2 step for the first timed leg against 75-80 ms when that start-up falls into the warm-up.)"""
        self.t_mark = time.perf_counter()

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
            self.lines.append((time.perf_counter(), line.strip()))

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
        timed = [ln for t, ln in self.lines if t >= self.t_mark]
        for ln in (timed if timed else [ln for _, ln in self.lines]):
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


def kernel_source_hash(workload):
    """sha256 (first 16 hex digits) of the dominant kernel's source + the shared device header: the ncu traffic record of
    a workload is only valid for the kernel it was captured from."""
    h = hashlib.sha256()
    for f in (WORKLOADS[workload]["kernel"], "sfc_device.cuh"):
        with open(os.path.join(ROOT, "sawfish_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_traffic(workload, n_pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed `ncu --set full`
    capture of this exact command (profiles/ncu_traffic.json).  None when no capture matches the workload, the batch
    size AND the current kernel source (records are stamped with kernel_source_hash)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    cur = kernel_source_hash(workload)
    for rec in json.load(open(p)):
        if rec["workload"] == workload and rec["pairs_per_gpu"] == n_pairs and rec.get("kernel_src_sha") == cur:
            return rec["dram_bytes_per_launch"]
    return None


def algorithmic_bytes(pairs, n_ops_words):
    """Bytes the path must move per batch: packed inputs (4 bit/base), 8 B score+status, run-length ops."""
    bases = int(pairs["pattern_len"].astype(np.int64).sum() + pairs["text_len"].astype(np.int64).sum())
    return bases // 2 + 8 * len(pairs) + 4 * int(n_ops_words)


def host_threads():
    # all host cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every
    # rank, which would turn the N>1 reference arm into a one-thread baseline.
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(workload, arena, pairs, budget_s, threads=0, keep_ops=False):
    """Times the CPU restatement (oracle) on a bounded sample of the batch: consecutive slices of the batch
    (in batch order, so the sample has the batch's own mix of matching / non-matching haplotypes) are aligned
    until ~budget_s seconds of wall time have been spent."""
    from oracle import oracle as O
    from helpers import rle_of_bytes
    cfg = WORKLOADS[workload]
    pen = O.CONSENSUS if cfg["preset"] == "consensus" else O.LARGE_INDEL
    threads = threads or host_threads()
    opairs = pairs.astype(O.PAIR_DTYPE)
    n = len(opairs)
    done, secs, cells, words, n_ops = 0, 0.0, 0, 0, 0
    chunk = max(2 * threads, 8)
    scores = np.zeros(n, np.int32)
    ops_rle = []
    while done < n and secs < budget_s:
        m = min(n - done, chunk)
        t0 = time.perf_counter()
        sc, status, stats, ops, off = O.align_batch(pen, arena, opairs[done:done + m], want_ops=True, mode=O.MODE_COMPACT, n_threads=threads)
        dt = time.perf_counter() - t0
        assert (status == 0).all()
        scores[done:done + m] = sc
        cells += int(stats["cells"].sum())
        words += int(stats["ext_words"].sum())
        n_ops += int(stats["n_ops"].sum())
        if keep_ops:
            for i in range(m):
                ops_rle.append(rle_of_bytes(ops[int(off[i]):int(off[i]) + int(stats["n_ops"][i])].tobytes()))
        done += m
        secs += dt
        if dt < budget_s / 8:
            chunk *= 2
    dp = float((opairs["pattern_len"][:done].astype(np.int64) * opairs["text_len"][:done].astype(np.int64)).sum())
    return {"pairs": done, "seconds": secs, "value": done / secs, "cores": threads, "cells": cells, "ext_words": words,
            "n_ops": n_ops, "gcups": dp / secs / 1e9, "scores": scores[:done], "ops_rle": ops_rle}


def config_dict(args, workload, reads, n_pairs_per_gpu, world, n_batches):
    """The `config` object — built the same way by both arms."""
    cfg = WORKLOADS[workload]
    return {"workload": cfg["desc"], "name": workload, "reads_per_gpu": reads, "pairs_per_gpu": int(n_pairs_per_gpu),
            "preset": cfg["preset"], "want_cigar": cfg["want_cigar"], "l2": L2_NOTE,
            "sharding": SHARD_NOTE if world > 1 else "single GPU: the whole batch",
            "batches": f"{n_batches} distinct batches cycled over the steps (seed {args.seed} + 1000 i)"}


def n_batches_for(args):
    return max(1, min(args.steps, args.batches))


def run_reference(args, rank, world):
    if rank != 0:
        return
    nb = n_batches_for(args)
    batches = [make_batch(args.workload, args.reads * world, args.seed + 1000 * b) for b in range(nb)]
    per_step_budget = max(2.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    info = None
    for it in range(args.warmup + args.steps):
        arena, pairs, _ = batches[it % nb]
        info = cpu_reference_run(args.workload, arena, pairs, per_step_budget)
        if it >= args.warmup:
            vals.append(info)
    tot_pairs = sum(v["pairs"] for v in vals)
    tot_s = sum(v["seconds"] for v in vals)
    val = tot_pairs / tot_s
    n_pairs_per_gpu = len(batches[0][1]) // world
    sample = f"first {info['pairs']} of {len(batches[0][1])} pairs of the step's batch (~{per_step_budget:.0f} s budget per step)"
    out = {"impl": "reference", "metric": "alignments_per_s", "value": val, "unit": "alignments/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(1, len(vals)),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": config_dict(args, args.workload, args.reads, n_pairs_per_gpu, world, nb),
           "gcups": sum(v["gcups"] * v["seconds"] for v in vals) / tot_s,
           "cpu_baseline": {"value": val, "unit": "alignments/s", "cores": info["cores"], "kind": "port", "sample": sample,
                            "note": "WFA2-semantics CPU restatement (oracle/), not the sawfish binary: no Rust toolchain / WFA2-lib here"},
           "e2e": {"value": val, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_REAL_STDOUT = None


def claim_stdout():
    """Rank 0 must print exactly ONE JSON line on stdout.  NCCL prints its version banner there (whenever NCCL_DEBUG is
    VERSION or higher) and other libraries may chat too, so file descriptor 1 is pointed at stderr for the whole run and
    the JSON line goes to a private duplicate of the original stdout.  (NCCL_DEBUG itself is left exactly as the caller
    set it: its output lands on stderr.)"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


class Dist:
    """torch.distributed plumbing: NCCL for the barrier / max-over-ranks of device times, a gloo group for the host-side
    gather of results (CPU tensors; there is no GPU collective on the data path)."""

    def __init__(self, rank, local_rank, world):
        import torch
        self.torch, self.rank, self.world = torch, rank, world
        self.dist = None
        self.gloo = None
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
            self.gloo = dist.new_group(backend="gloo")

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.dist:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def run_leg(args, D, ctx, workload, reads, steps, warmup, cpu_budget, local_rank, sample_clocks):
    """One workload on this rank's GPU: warm-up, device-resident leg, end-to-end leg, verification, CPU baseline.
    Returns the JSON fields of the leg (rank 0) or None."""
    import torch
    from sawfish_b200 import sharding, synth
    from sawfish_b200.batch import CONSENSUS_PENALTIES, LARGE_INDEL_PENALTIES
    from helpers import rle_of_bytes  # noqa: F401  (the oracle side of the verification uses it)
    rank, world = D.rank, D.world
    cfg = WORKLOADS[workload]
    pen = CONSENSUS_PENALTIES if cfg["preset"] == "consensus" else LARGE_INDEL_PENALTIES
    want_cigar = cfg["want_cigar"]
    nb = max(1, min(steps, args.batches))

    # ---- the global batches (identical on every rank) and this rank's shard of each ----
    t_gen = time.perf_counter()
    glob = [make_batch(workload, reads * world, args.seed + 1000 * b) for b in range(nb)]
    t_gen = time.perf_counter() - t_gen

    def cut(b):
        arena, pairs, groups = glob[b]
        if world == 1:
            return np.arange(len(pairs)), arena, pairs
        mine, _plan = sharding.shard_pairs(pen, pairs, groups, world, rank)
        la, lp = compact_shard(arena, pairs, mine)
        return mine, la, lp

    shards = [cut(b) for b in range(nb)]
    max_arena = max(len(s[1]) for s in shards)
    max_pairs = max(len(s[2]) for s in shards)
    # pinned host staging (one buffer set per batch, filled once: the step's inputs "arrive" in pinned memory)
    pins = []
    for mine, la, lp in shards:
        a = torch.from_numpy(la).pin_memory()
        p = torch.from_numpy(lp.view(np.uint8).copy()).pin_memory()
        pins.append((a, p, len(lp)))
    scores_pin = torch.empty(max_pairs, dtype=torch.int32).pin_memory()
    status_pin = torch.empty(max_pairs, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def upload(b):
        a, p, n = pins[b]
        ctx.upload_ptr(a.data_ptr(), a.numel(), p.data_ptr(), n)
        return n

    def e2e_step(b):
        n = upload(b)
        ctx.run(pen, want_cigar)
        return ctx.download(out=(scores_pin.numpy()[:n], status_pin.numpy()[:n]))

    def gather(b, res):
        """Host-side gather on rank 0 in original pair order: scores always, run-length ops when the workload has them."""
        scores, status, ops, ops_off = res
        mine = shards[b][0]
        if world == 1:
            return scores, ops, ops_off
        payload = (mine, np.asarray(scores, np.int32), None if ops is None else np.asarray(ops), None if ops_off is None else np.asarray(ops_off))
        got = [None] * world if rank == 0 else None
        D.dist.gather_object(payload, got, dst=0, group=D.gloo)
        if rank != 0:
            return None, None, None
        return merge_rank_results(got, len(glob[b][1]), want_cigar)

    # ---- warm-up (also allocates device workspace) ----
    # At least W (>= 3) steps AND at least ~2.5 s: the first GPU process on a fresh box runs its first few hundred
    # milliseconds measurably slower, and the driver's bench is exactly that process.
    sampler = None
    if sample_clocks:
        sampler = ClockSampler(local_rank)
        sampler.start()
    n_warm = 0
    t_warm = time.perf_counter()
    res0 = None
    while n_warm < max(3, warmup) or (time.perf_counter() - t_warm < 2.5 and n_warm < 64):
        r = e2e_step(n_warm % nb)
        assert (r[1] == 0).all(), f"pair status != 0: {np.unique(r[1])}"
        if n_warm % nb == 0 and res0 is None:
            res0 = (r[0].copy(), r[1].copy(), None if r[2] is None else r[2].copy(), None if r[3] is None else r[3].copy())
        n_warm += 1
    g0 = gather(0, res0)   # the gathered results of batch 0: what the verification below checks

    flush.fill_(1)  # first use of the flush kernel (module load) belongs to the warm-up as well
    torch.cuda.synchronize()
    if sampler:
        sampler.mark()
    # ---- device-resident throughput: kernels only ----
    D.barrier()
    kern_ms, launches, cells, cells0, retries = 0.0, 0, 0, 0, 0
    step_ms, e2e_step_ms = [], []
    n_ops_words = 0
    alg_bytes = 0
    t0 = time.perf_counter()
    for it in range(steps):
        b = it % nb
        upload(b)
        flush.fill_(1)  # L2 flush between timed iterations
        torch.cuda.synchronize()
        ctx.run(pen, want_cigar)
        torch.cuda.synchronize()
        # kernel_ms is measured by CUDA events on the library's own stream around the launches of this run
        r = ctx.download(out=(scores_pin.numpy()[:pins[b][2]], status_pin.numpy()[:pins[b][2]]))
        s = ctx.stats()
        kern_ms += s["kernel_ms"]
        step_ms.append(round(float(s["kernel_ms"]), 2))
        launches += s["launches"] - 1  # minus the pack kernel, which belongs to upload
        cells += s["cells"]
        retries += s["retries"]
        if b == 0:
            cells0 = s["cells"]
        n_ops_words = 0 if r[2] is None else len(r[2])
        alg_bytes += algorithmic_bytes(shards[b][2], n_ops_words)
    D.barrier()
    wall_resident = time.perf_counter() - t0
    # ---- end to end through the C ABI from pinned host buffers ----
    e2e_wall = 0.0
    e2e_dev_ms = 0.0
    plan_ms = 0.0
    h2d_b = d2h_b = 0
    for it in range(steps):
        b = it % nb
        D.barrier()
        t0 = time.perf_counter()
        if world > 1:
            # the sharder is part of the step: work estimate + LPT plan (C ABI) and the compaction of this rank's shard into
            # its pinned staging buffers
            tp = time.perf_counter()
            arena, pairs, groups = glob[b]
            mine, _ = sharding.shard_pairs(pen, pairs, groups, world, rank)
            a, p, n = pins[b]
            la, lp = compact_shard(arena, pairs, mine, out=a.numpy())
            p.numpy()[:lp.nbytes] = lp.view(np.uint8)
            plan_ms += (time.perf_counter() - tp) * 1e3
        r = e2e_step(b)
        s = ctx.stats()
        gather(b, r)
        D.barrier()
        e2e_wall += time.perf_counter() - t0
        e2e_dev_ms += s["h2d_ms"] + s["pack_ms"] + s["kernel_ms"] + s["d2h_ms"]
        e2e_step_ms.append(round(float(s["kernel_ms"]), 2))
        h2d_b += int(s["h2d_bytes"]); d2h_b += int(s["d2h_bytes"])
        retries += s["retries"]
    clocks = sampler.stop() if sampler else None
    st = ctx.stats()

    # max over ranks (device time), whole-job aggregate
    kern_ms_max, e2e_wall_ms_max, e2e_dev_ms_max, plan_ms_max = D.reduce([kern_ms, e2e_wall * 1e3, e2e_dev_ms, plan_ms], "MAX")
    pairs_per_cycle = [float(len(glob[b][1])) for b in range(nb)]
    all_pairs = sum(pairs_per_cycle[it % nb] for it in range(steps))       # global pairs aligned in the timed steps
    all_cells_dp = sum(float(synth.gcups_cells(glob[it % nb][1])) for it in range(steps))
    all_launches, all_cells, all_alg_bytes, all_retries = D.reduce([float(launches), float(cells), float(alg_bytes), float(retries)], "SUM")  # retries: both legs
    (kern_ms_min,) = D.reduce([kern_ms], "MIN")
    if rank != 0:
        return None
    ms_per_step = kern_ms_max / steps
    value = all_pairs / (kern_ms_max * 1e-3)
    e2e_value = all_pairs / (e2e_wall_ms_max * 1e-3)
    peak, peak_src = measured_peaks()
    achieved = all_alg_bytes / world / (kern_ms_max * 1e-3) / 1e9   # per GPU, against one GPU's peak
    int_peak_gops, _ = ctx.measure_int_peak()
    out = {
        "metric": "alignments_per_s", "value": value, "unit": "alignments/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "warmup_done": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        # the arithmetic the dominant kernel computes in: K2 v2 works on 16-bit offsets packed two to a word, K1 on 32-bit cells
        "vs_baseline": None, "dtype": "int16" if cfg["kernel"] == "sfc_k_align2.cuh" else "int32", "data": "synthetic",
        "config": config_dict(args, workload, reads, len(glob[0][1]) // world, world, nb),
        "gcups": all_cells_dp / (kern_ms_max * 1e-3) / 1e9,
        "wavefront_cells_per_step": int(all_cells / steps),
        "gpu_launches": int(all_launches),
        "e2e": {"value": e2e_value, "unit": "alignments/s", "h2d_bytes_per_step": h2d_b // steps,
                "d2h_bytes_per_step": d2h_b // steps, "ms_per_step_wall": e2e_wall_ms_max / steps,
                "ms_per_step_device": e2e_dev_ms_max / steps,
                "shard_plan_and_compaction_ms_per_step": plan_ms_max / steps if world > 1 else 0.0,
                "note": "rank 0's bytes; at N > 1 the timed step also holds the LPT shard plan, the shard compaction and the host-side gather"},
        "kernel_split": {"n_score_kernel_pairs": st["n_score_kernel"], "n_align_kernel_pairs": st["n_align_kernel"],
                         "retries": int(all_retries), "pack_ms": st["pack_ms"], "h2d_ms": st["h2d_ms"], "d2h_ms": st["d2h_ms"]},
        "wall_ms_per_step_resident": wall_resident * 1e3 / steps,
        "shard_balance": {"kernel_ms_min_over_ranks": kern_ms_min / steps, "kernel_ms_max_over_ranks": kern_ms_max / steps},
        "kernel_ms_per_step_rank0": {"resident": step_ms, "e2e": e2e_step_ms},
        "batch_generation_s": t_gen,
    }
    if clocks is not None:
        out["clocks"] = clocks
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": ncu_traffic(workload, len(shards[0][2])) if world == 1 else None, "peak_source": peak_src,
                       "kernel_src_sha": kernel_source_hash(workload),
                       "note": "algorithmic bytes = 4-bit packed inputs + scores/status + run-length ops; the path is integer-ALU/latency bound, see roofline_int; "
                               "traffic = dram bytes per launch of the dominant kernel from the committed ncu capture, null unless it was taken from this kernel source"}
    # ---- verification + CPU baseline on a bounded sample of batch 0 (global pair order) ----
    arena0, pairs0, _ = glob[0]
    g_scores, g_ops, g_off = g0
    budget = cpu_budget if (world == 1 and not args.no_cpu_baseline) else min(cpu_budget, 4.0)
    cb = cpu_reference_run(workload, arena0, pairs0, budget, keep_ops=want_cigar)
    m = cb["pairs"]
    mism = int((cb["scores"] != g_scores[:m]).sum())
    ops_mism = 0
    if want_cigar:
        for i in range(m):
            got = g_ops[int(g_off[i]):int(g_off[i + 1])]
            want = cb["ops_rle"][i]
            if len(got) != len(want) or not (np.asarray(got) == want).all():
                ops_mism += 1
    out["verified"] = {"pairs": m, "mismatches": mism + ops_mism, "score_mismatches": mism,
                       "ops_mismatches": ops_mism if want_cigar else None,
                       "what": "scores" + (" and run-length ops" if want_cigar else "") + " of the first pairs of batch 0 (global order, after the gather) vs the CPU restatement"}
    # integer work from the oracle's own counters: C wavefront cells, W 16-base extend words
    aff = cfg["preset"] == "large_indel"
    per_c, per_w = (20, 6) if aff else (6, 6)
    samp_ops = per_c * cb["cells"] + per_w * cb["ext_words"] + (4 * cb["n_ops"] if want_cigar else 0)
    if m == len(pairs0) and world == 1 and nb == 1:
        int_ops_step, how = float(samp_ops), "oracle counters of the whole batch"
    else:
        # the oracle's ops per wavefront cell on the sample, times the device's own cell counter of the timed steps
        int_ops_step = samp_ops / max(1, cb["cells"]) * (all_cells / steps)
        how = "oracle (20C+6W | 6C+6W, +4 per emitted op) per cell on the verified sample x device cell counter of the timed steps"
    achieved_int = int_ops_step / world / (ms_per_step * 1e-3) / 1e9
    out["roofline_int"] = {"bound": "int32-alu", "achieved": achieved_int, "peak": int_peak_gops, "unit": "Gop/s",
                           "frac": achieved_int / int_peak_gops, "ops_per_cell": samp_ops / max(1, cb["cells"]), "how": how,
                           "oracle_cells_sample": cb["cells"], "oracle_ext_words_sample": cb["ext_words"],
                           "device_cells_batch0": int(cells0) if world == 1 else None,
                           "peak_source": "IADD/LOP3 micro-benchmark run in this process (sfc_measure_int_peak)"}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = {"value": cb["value"], "unit": "alignments/s", "cores": cb["cores"], "kind": "port",
                               "sample": f"first {cb['pairs']} of {len(pairs0)} pairs of batch 0, {cb['seconds']:.1f} s",
                               "gcups": cb["gcups"],
                               "note": "WFA2-semantics CPU restatement (oracle/, OpenMP over pairs, fresh aligner state per pair)"}
    if mism + ops_mism and not args.no_verify_fail:
        out["verified"]["FAILED"] = True
    return out


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="reads (or clusters / SV groups) per GPU per step")
    ap.add_argument("--seed", type=int, default=0x5AF15 + 2)
    ap.add_argument("--batches", type=int, default=3, help="distinct batches cycled over the steps")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU baseline work (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the timed CPU leg (a short verification sample still runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config2b / config4 legs of the default line")
    ap.add_argument("--no-verify-fail", action="store_true", help="report mismatches in `verified` without a non-zero exit")
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
    from sawfish_b200.batch import AlignContext

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    D = Dist(rank, local_rank, world)
    ctx = AlignContext(local_rank)
    out = run_leg(args, D, ctx, args.workload, args.reads, args.steps, args.warmup, args.cpu_budget, local_rank, True)
    failed = bool(out and out["verified"].get("FAILED"))
    if world == 1 and args.workload == "config2" and not args.no_secondary:
        sec = {}
        for wl, reads in SECONDARY.items():
            leg = run_leg(args, D, ctx, wl, reads, max(2, min(args.steps, 3)), 3, min(args.cpu_budget, 6.0), local_rank, False)
            failed = failed or bool(leg["verified"].get("FAILED"))
            sec[wl] = {k: leg[k] for k in ("value", "unit", "ms_per_step", "gcups", "config", "e2e", "roofline_int", "verified",
                                           "gpu_launches", "wavefront_cells_per_step") if k in leg}
            if "cpu_baseline" in leg:
                sec[wl]["cpu_baseline"] = leg["cpu_baseline"]
        out["secondary"] = sec
    if rank == 0:
        emit(out)
    ctx.close()
    D.close()
    if failed:
        raise SystemExit("bench.py: GPU results differ from the CPU restatement on the verified sample")


if __name__ == "__main__":
    main()
