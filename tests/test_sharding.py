"""N>1 host-side logic on CPU: gloo, world_size 2.  The per-shard aligner here is the oracle (checker), so this
tests partition + gather + ordering, not the CUDA kernels."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from sawfish_b200 import synth
        from sawfish_b200.batch import CONSENSUS_PENALTIES
        from sawfish_b200.sharding import gather_results, shard_pairs
        arena, pairs, _ = synth.config2b_score_sv_flanks(60, seed=99)
        group_ids = np.arange(len(pairs)) // 4          # 4 pairs (2 reads x 2 alleles) per "SV group"
        mine, plan = shard_pairs(CONSENSUS_PENALTIES, pairs, group_ids, world, rank)
        assert (plan[group_ids[mine]] == rank).all()
        sub = np.ascontiguousarray(pairs[mine]).astype(O.PAIR_DTYPE)
        scores, status, *_ = O.align_batch(O.CONSENSUS, arena, sub, want_ops=False, n_threads=1)
        assert (status == 0).all()
        full = gather_results(mine, scores, len(pairs), dst=0)
        if rank == 0:
            ref, st, *_ = O.align_batch(O.CONSENSUS, arena, pairs.astype(O.PAIR_DTYPE), want_ops=False, n_threads=2)
            assert (full == ref).all()
            ret.put(("ok", int(len(mine)), int(len(pairs))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    tag, n_mine, n_all = ret.get(timeout=5)
    assert tag == "ok" and 0 < n_mine < n_all


def test_shard_plan_balanced_and_deterministic():
    from sawfish_b200.sharding import plan_shards
    rng = np.random.default_rng(0)
    w = (rng.pareto(1.5, 500) * 1000 + 1).astype(np.uint64)
    for shards in (1, 2, 4, 8):
        plan = plan_shards(w, shards)
        assert plan.max() < shards and (plan_shards(w, shards) == plan).all()
        loads = np.bincount(plan, weights=w.astype(np.float64), minlength=shards)
        assert loads.max() <= loads.mean() + float(w.max())   # LPT bound


def test_estimate_pair_work_monotone():
    from sawfish_b200.batch import LARGE_INDEL_PENALTIES, PAIR_DTYPE
    from sawfish_b200.sharding import estimate_pair_work
    rows = [(0, 0, 5000, 5000, 2000, 2000, 2000, 2000, 1), (0, 0, 5000, 9000, 2000, 2000, 2000, 2000, 1),
            (0, 0, 500, 500, 200, 200, 200, 200, 1)]
    w = estimate_pair_work(LARGE_INDEL_PENALTIES, np.array(rows, dtype=PAIR_DTYPE))
    assert w[1] > w[0] > w[2] > 0


def test_bench_shard_cut_covers_every_pair_once_and_compacts_exactly():
    """What bench.py does per rank at N > 1: the LPT plan over SV groups (identical on every rank), this rank's pairs, and a
    compacted arena holding exactly their sequences — every pair lands on exactly one rank with its bytes unchanged."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from sawfish_b200.batch import LARGE_INDEL_PENALTIES
    from sawfish_b200.sharding import shard_pairs
    arena, pairs, groups = bench.make_batch("config2", 24, 11)
    for world in (2, 4, 8):
        seen = np.zeros(len(pairs), int)
        for rank in range(world):
            mine, plan = shard_pairs(LARGE_INDEL_PENALTIES, pairs, groups, world, rank)
            seen[mine] += 1
            assert len(np.unique(plan[groups[mine]])) <= 1          # whole groups only
            la, lp = bench.compact_shard(arena, pairs, mine)
            assert len(la) <= len(arena)
            for i, j in enumerate(mine[:40]):
                a, b = pairs[j], lp[i]
                for off, ln in (("pattern_off", "pattern_len"), ("text_off", "text_len")):
                    assert b[ln] == a[ln]
                    assert (arena[int(a[off]):int(a[off]) + int(a[ln])] == la[int(b[off]):int(b[off]) + int(b[ln])]).all()
        assert (seen == 1).all()
    arena4, pairs4, groups4 = bench.make_batch("config4", 3, 5)
    assert len(groups4) == len(pairs4) and groups4.max() == 2


def test_bench_merge_rank_results_puts_ragged_ops_back_in_global_order():
    """bench.py's rank-0 gather (scores + ragged run-length ops of every rank -> global pair order), against a per-pair
    reference; a pair on two ranks or on none must be refused."""
    import bench
    rng = np.random.default_rng(11)
    n = 300
    lens = rng.integers(0, 7, n)
    truth = [rng.integers(0, 1 << 20, l).astype(np.uint32) for l in lens]
    scores = rng.integers(-500, 0, n).astype(np.int32)
    owner = rng.integers(0, 3, n)
    parts = []
    for r in range(3):
        idx = np.flatnonzero(owner == r)
        off = np.concatenate([[0], np.cumsum(lens[idx])]).astype(np.uint64)
        op = np.concatenate([truth[i] for i in idx]) if len(idx) else np.zeros(0, np.uint32)
        parts.append((idx, scores[idx], op, off))
    g_scores, g_ops, g_off = bench.merge_rank_results(parts, n, True)
    assert (g_scores == scores).all()
    assert all((g_ops[int(g_off[i]):int(g_off[i + 1])] == truth[i]).all() for i in range(n))
    s2, o2, f2 = bench.merge_rank_results(parts, n, False)
    assert (s2 == scores).all() and o2 is None and f2 is None
    with pytest.raises(AssertionError):
        bench.merge_rank_results(parts + [parts[0]], n, True)   # assigned twice
    with pytest.raises(AssertionError):
        bench.merge_rank_results(parts[:2], n, True)            # somebody's pairs are missing
