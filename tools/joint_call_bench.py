#!/usr/bin/env python
"""joint-call wall time on a synthetic multi-sample directory set (BASELINE.json metric "joint-call wall-time";
configs[3] in miniature), GPU path next to the CPU restatement, with the reference's counter names.

What is timed is the part of `sawfish joint-call` that sits on the alignment hot path (reference src/joint_call/mod.rs:25-118):

  load       per-sample discover directories: contig.alignment.bam + candidate.sv.bcf          get_refined_svs.rs:176-195,554-601
  merge      duplicate-haplotype pooling over all samples: the rust-bio merge aligner             merge_pools.rs:214-441
             (GPU: one K5 batch;  CPU: the oracle's restatement of the same aligner)
  scoring    per SV group x sample: read flanks x allele flanks -> one alignment batch ->        score_sv/mod.rs:758-905,970-1063,
             normalised scores -> best allele per read -> AD -> GT / GQ / PL -> QUAL             1130-1369,1422-1549
             (GPU: K1;  CPU: the oracle on all host threads)  == ScoreStats::total_scoring_time_secs (run_stats.rs:83)

The reads come from a synthetic "BAM" (pre-generated read flanks per locus and sample, not timed): the BAM fetch itself
is htslib I/O and out of scope.  Both arms run the same host code (Python mirrors); only the aligner differs,
and the genotypes of the two arms must be identical.

usage: python tools/joint_call_bench.py [--loci 96] [--samples 8] [--depth 30] [--seed 1] [--no-cpu]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOCUS_STRIDE = 20_000
WIN = 4000          # reference window per locus
POS = 2000          # breakpoint position inside the window
CFLANK = 250        # contig flank either side of the insertion


def make_cohort(n_loci, n_samples, depth, seed):
    """Truth + the synthetic inputs of joint-call: per-sample discover output and per (locus, sample) read flanks."""
    from sawfish_b200 import synth
    from sawfish_b200 import score_sv as SS
    rng = np.random.default_rng(seed)
    half = SS.READ_SUPPORT_FLANK_SIZE // 2
    loci = []
    for li in range(n_loci):
        ref = synth.random_seq(rng, WIN)
        ins = synth.random_seq(rng, synth.log_uniform_int(rng, 40, 400))
        alt = np.concatenate([ref[:POS], ins, ref[POS:]])
        gts = rng.choice([0, 1, 2], size=n_samples, p=[0.45, 0.4, 0.15])
        if not (gts > 0).any():
            gts[int(rng.integers(0, n_samples))] = 1
        loci.append(dict(ref=ref, ins=ins, alt=alt, gts=gts))
    # reads: per (locus, sample) `depth` reads, each with a flank at both breakends of its own haplotype
    reads = {}
    for li, L in enumerate(loci):
        for si in range(n_samples):
            gt = int(L["gts"][si])
            lst = []
            for ri in range(depth):
                from_alt = rng.random() < (0.0, 0.5, 1.0)[gt]
                hap = L["alt"] if from_alt else L["ref"]
                bps = (POS, POS + len(L["ins"])) if from_alt else (POS, POS)
                flanks = []
                for be in (0, 1):
                    c = bps[be]
                    st = int(rng.integers(0, 120)) if rng.random() < 0.15 else 0
                    et = int(rng.integers(0, 120)) if rng.random() < 0.15 else 0
                    flanks.append((synth.add_errors(rng, hap[c - half + st:c + half - et], 0.001), st, et))
                lst.append((f"read{ri:04d}", flanks))
            reads[(li, si)] = lst
    return loci, reads


def write_discover_dirs(loci, n_samples, root, seed):
    from sawfish_b200 import candidate_sv as CS
    from sawfish_b200 import contig_bam as CB
    rng = np.random.default_rng(seed + 77)
    chroms = [("chr1", LOCUS_STRIDE * (len(loci) + 2))]
    acgt = np.frombuffer(b"ACGT", np.uint8)
    types = {"SVTYPE": "String", "END": "Integer", "SVLEN": "Integer", "CONTIG_POS": "Integer"}
    dirs = []
    for s in range(n_samples):
        contigs, recs = [], []
        ci = 0
        for li, L in enumerate(loci):
            if L["gts"][s] == 0:
                continue
            n_ins = len(L["ins"])
            seq = L["alt"][POS - CFLANK:POS + n_ins + CFLANK].copy()
            for _e in range(int(rng.integers(0, 3))):       # assembly noise: the reason pooling needs an aligner at all
                p = int(rng.integers(0, len(seq)))
                seq[p] = acgt[int(rng.integers(0, 4))]
            pos = li * LOCUS_STRIDE + POS - CFLANK            # leftmost aligned reference base of the contig
            segs = [CB.ContigAlignmentSegment(0, pos, [("=", CFLANK), ("I", n_ins), ("=", CFLANK)])]
            contigs.append(CB.ContigAlignmentInfo(s, ci, 0, seq.tobytes(), 0, segs, int(rng.integers(3, 30)),
                                                  (int(rng.integers(0, 20)), len(seq) - int(rng.integers(0, 20)))))
            anchor = pos + CFLANK - 1
            ref_base = chr(seq[CFLANK - 1]).encode()
            recs.append(CS.CandidateRecord(0, anchor, f"sawfish:{s}:{ci}:0:0", [ref_base, ref_base + seq[CFLANK:CFLANK + n_ins].tobytes()],
                                           {"SVTYPE": ["INS"], "END": [anchor + 1], "SVLEN": [n_ins], "CONTIG_POS": [CFLANK - 1]}))
            ci += 1
        d = os.path.join(root, f"discover{s}")
        os.makedirs(d)
        CB.write_contig_alignments(os.path.join(d, CB.CONTIG_ALIGNMENT_FILENAME), chroms, contigs)
        CS.write_candidate_bcf(os.path.join(d, CS.CANDIDATE_SV_FILENAME), chroms, recs, types)
        dirs.append(d)
    return dirs, chroms


def joint_call(dirs, chroms, loci, reads, n_samples, merge_score_fn, align_fn):
    """The timed pipeline.  Returns (genotype table, counters)."""
    from sawfish_b200 import discover_dir as DD
    from sawfish_b200 import merge_pools as MP
    from sawfish_b200 import score_sv as SS
    T = {}
    t0 = time.perf_counter()
    l2i = {n: i for i, (n, _l) in enumerate(chroms)}
    samples = [DD.load_discover_dir(d, l2i) for d in dirs]
    T["load_time_secs"] = time.perf_counter() - t0
    # ---- merge: duplicate haplotypes across samples ----
    t0 = time.perf_counter()
    haps, sample_lists = DD.merge_haplotypes_of_samples(samples)
    pools, mstats = MP.get_haplotype_merge_pools(haps, sample_lists, merge_score_fn)
    groups = [p for p in pools if p]
    T["merge_time_secs"] = time.perf_counter() - t0
    # ---- scoring: every merged SV group against every sample's reads ----
    t0 = time.perf_counter()
    half = SS.READ_SUPPORT_FLANK_SIZE // 2
    b = SS.PairBatchBuilder()
    tags = []
    for gi, pool in enumerate(groups):
        rep = haps[pool[0]]
        bp = rep.breakpoints[0]
        li = bp.breakend1.start // LOCUS_STRIDE
        ref = loci[li]["ref"]
        contig = np.frombuffer(rep.contig_seq, np.uint8)
        n_ins = bp.insert_size
        alt_centers, ref_centers = (CFLANK, CFLANK + n_ins), (POS, POS)
        for si in range(n_samples):
            tag = gi * n_samples + si
            tags.append(tag)
            for qname, flanks in reads[(li, si)]:
                for be in (SS.BREAKEND1, SS.BREAKEND2):
                    seq, st, et = flanks[be]
                    rc, ac = ref_centers[be], alt_centers[be]
                    alt_flank = contig[max(0, ac - half):ac + half]
                    b.add_read_flank(tag, qname, be, SS.STANDARD, SS.ReadFlank(seq, st, et), ref[rc - half:rc + half], alt_flank, [])
    t_enum = time.perf_counter() - t0
    arena, pairs = b.finish()
    ta = time.perf_counter()
    scores = align_fn(arena, pairs)
    t_align = time.perf_counter() - ta
    tg = time.perf_counter()
    locus = SS.scatter_scores(b.slots, scores, {t: [] for t in tags})
    table = {}
    for gi in range(len(groups)):
        ads = []
        for si in range(n_samples):
            tag = gi * n_samples + si
            ads.append(SS.set_allele_depth_counts(locus.get(tag, {}), sv_haplotype_index=0, overlapping_haplotype_index=None,
                                                  is_alt_on_top_haplotype_rank=True, on_multiple_sample_haplotypes=False, read_to_hap={}))
        sq, qual = SS.get_sv_qualities_from_allele_counts(ads)
        table[gi] = ([(int(q.gt) if q.gt is not None else -1, q.gq, tuple(ad)) for q, ad in zip(sq, ads)], qual)
    T["total_scoring_time_secs"] = time.perf_counter() - t0
    T["scoring_enumeration_secs"] = t_enum
    T["scoring_alignment_secs"] = t_align
    T["scoring_genotype_secs"] = time.perf_counter() - tg
    T["merge_alignments"] = mstats["alignments_batched"]
    T["sv_groups_after_merge"] = len(groups)
    T["score_pairs"] = int(len(pairs))
    T["wall_time_secs"] = T["load_time_secs"] + T["merge_time_secs"] + T["total_scoring_time_secs"]
    return table, T, groups, haps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--loci", type=int, default=96)
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from oracle import oracle as O
    from sawfish_b200 import merge_pools as MP
    from sawfish_b200 import score_sv as SS
    from sawfish_b200.batch import AlignContext
    loci, reads = make_cohort(args.loci, args.samples, args.depth, args.seed)
    out = {"workload": f"synthetic joint-call: {args.loci} insertion loci x {args.samples} samples x depth {args.depth}", "arms": {}}
    with tempfile.TemporaryDirectory() as root:
        dirs, chroms = write_discover_dirs(loci, args.samples, root, args.seed)
        ctx = AlignContext(0)
        gpu_align = SS.gpu_align_fn(ctx)
        gpu_merge = MP.gpu_merge_score_fn(ctx)
        joint_call(dirs, chroms, loci, reads, args.samples, gpu_merge, gpu_align)          # warm-up (allocations, first launches)
        g_table, g_T, groups, haps = joint_call(dirs, chroms, loci, reads, args.samples, gpu_merge, gpu_align)
        out["arms"]["gpu"] = g_T
        if not args.no_cpu:
            threads = len(os.sched_getaffinity(0))
            merge_sc = O.bio_scoring(0, -2, 1, -3, yclip_prefix=0, yclip_suffix=0)

            def cpu_merge(batch):
                return O.bio_custom_score_batch(merge_sc, batch, threads) if hasattr(O, "bio_custom_score_batch") else \
                    [O.bio_custom_align(merge_sc, q, bb)[0] for bb, q in batch]

            def cpu_align(arena, pairs):
                sc, st, *_ = O.align_batch(O.CONSENSUS, arena, pairs.astype(O.PAIR_DTYPE), want_ops=False, n_threads=threads)
                assert (st == 0).all()
                return sc
            c_table, c_T, _, _ = joint_call(dirs, chroms, loci, reads, args.samples, cpu_merge, cpu_align)
            c_T["host_threads"] = threads
            out["arms"]["cpu"] = c_T
            out["genotypes_identical"] = g_table == c_table
            out["speedup_total_scoring_time"] = c_T["total_scoring_time_secs"] / g_T["total_scoring_time_secs"]
            out["speedup_scoring_alignment"] = c_T["scoring_alignment_secs"] / g_T["scoring_alignment_secs"]
            out["speedup_wall_time"] = c_T["wall_time_secs"] / g_T["wall_time_secs"]
        # accuracy against the planted genotypes (information only)
        ok = tot = 0
        for gi, pool in enumerate(groups):
            li = haps[pool[0]].breakpoints[0].breakend1.start // LOCUS_STRIDE
            for si in range(args.samples):
                tot += 1
                ok += int(g_table[gi][0][si][0] == int(loci[li]["gts"][si]))
        out["genotype_concordance_with_truth"] = ok / max(1, tot)
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
