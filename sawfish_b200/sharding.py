"""Multi-GPU sharding of alignment batches: one process per GPU, no data-path collective.

The unit of sharding is the SV group / breakpoint cluster — the reference's own rayon task granularity
(src/joint_call/joint_call_all_samples.rs:127-142, src/refine_sv/mod.rs:861-887) — because
add_guest_haplotypes (src/score_sv/mod.rs:2293) needs all scores of a group together.  Groups are assigned to
ranks by greedy longest-processing-time over estimated wavefront work (C ABI: sfc_estimate_pair_work,
sfc_shard_plan); every rank aligns its own groups on its own GPU and the integer results are gathered on the
host in original pair order (the reference gathers through an mpsc channel:
src/joint_call/joint_call_all_samples.rs:144-149).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import Penalties, SfcError


def estimate_pair_work(pen: Penalties, pairs: np.ndarray) -> np.ndarray:
    L = _lib.lib()
    work = np.zeros(len(pairs), np.uint64)
    rc = L.sfc_estimate_pair_work(C.byref(pen), pairs.ctypes.data, len(pairs), work.ctypes.data)
    if rc != 0:
        raise SfcError(rc, "sfc_estimate_pair_work")
    return work


def plan_shards(group_work: np.ndarray, n_shards: int) -> np.ndarray:
    L = _lib.lib()
    group_work = np.ascontiguousarray(group_work, np.uint64)
    out = np.zeros(len(group_work), np.uint32)
    rc = L.sfc_shard_plan(group_work.ctypes.data, len(group_work), n_shards, out.ctypes.data)
    if rc != 0:
        raise SfcError(rc, "sfc_shard_plan")
    return out


def shard_pairs(pen: Penalties, pairs: np.ndarray, group_ids: np.ndarray, world: int, rank: int):
    """Returns (indices of this rank's pairs in the original batch, shard_of_group)."""
    group_ids = np.asarray(group_ids, np.int64)
    n_groups = int(group_ids.max()) + 1 if len(group_ids) else 0
    work = estimate_pair_work(pen, pairs)
    gwork = np.bincount(group_ids, weights=work.astype(np.float64), minlength=n_groups).astype(np.uint64)
    plan = plan_shards(gwork, world)
    mine = np.flatnonzero(plan[group_ids] == rank)
    return mine, plan


def gather_results(local_idx: np.ndarray, local_scores: np.ndarray, n_total: int, dst: int = 0, group=None) -> Optional[np.ndarray]:
    """Host-side gather of per-pair int32 results to rank `dst`, restored to original pair order.
    Uses CPU tensors (gloo) — there is no GPU collective on this path."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = np.zeros(n_total, np.int32)
        out[local_idx] = local_scores
        return out
    payload = (np.asarray(local_idx, np.int64), np.asarray(local_scores, np.int32))
    gathered = [None] * dist.get_world_size(group) if dist.get_rank(group) == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if dist.get_rank(group) != dst:
        return None
    out = np.zeros(n_total, np.int32)
    seen = np.zeros(n_total, bool)
    for idx, sc in gathered:
        assert not seen[idx].any(), "a pair was assigned to two ranks"
        out[idx] = sc
        seen[idx] = True
    assert seen.all(), "a pair was assigned to no rank"
    return out
