"""Batch front-end over the C ABI: one AlignContext per (process, GPU).

This is what refine_sv / score_sv batching code calls instead of looping over
wfa2_utils::PairwiseAligner::align (reference src/score_sv/mod.rs:837-901,
src/refine_sv/align/mod.rs:439-451,2027-2049).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import BatchStats, Penalties, SfcError

PAIR_DTYPE = np.dtype([("pattern_off", "<u8"), ("text_off", "<u8"), ("pattern_len", "<u4"), ("text_len", "<u4"),
                       ("pattern_begin_free", "<i4"), ("pattern_end_free", "<i4"), ("text_begin_free", "<i4"),
                       ("text_end_free", "<i4"), ("flags", "<u4")], align=True)
assert PAIR_DTYPE.itemsize == 48
PAIR_REVERSE = 1

# reference src/wfa2_utils.rs:183-189 / :159-168
CONSENSUS_PENALTIES = Penalties(0, -1, 3, 0, 2, 0, 0)
LARGE_INDEL_PENALTIES = Penalties(1, -1, 3, 1, 2, 60, 0)


def sawfish_clip(plen: int, tlen: int) -> int:
    """clip = min(floor(0.4|p|), floor(0.4|t|)) with f64 -> i32 truncation (src/wfa2_utils.rs:111-115)."""
    return min(int(float(plen) * 0.4), int(float(tlen) * 0.4))


class AlignContext:
    def __init__(self, device: int = 0):
        L = _lib.lib()
        h = C.c_void_p()
        rc = L.sfc_create(device, C.byref(h))
        if rc != 0:
            raise SfcError(rc, "sfc_create failed (a B200 / sm_100 device is required; there is no CPU fallback)")
        self._h = h
        self._L = L
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.sfc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise SfcError(rc, (self._L.sfc_last_error(self._h) or b"").decode())

    # ---- staged API ----
    def upload(self, arena: np.ndarray, pairs: np.ndarray):
        assert arena.dtype == np.uint8 and arena.flags.c_contiguous
        assert pairs.dtype == PAIR_DTYPE and pairs.flags.c_contiguous
        self._keep = (arena, pairs)
        self._n = len(pairs)
        self._check(self._L.sfc_batch_upload(self._h, arena.ctypes.data, arena.size, pairs.ctypes.data, len(pairs)))

    def upload_ptr(self, arena_ptr: int, arena_len: int, pairs_ptr: int, n_pairs: int):
        """Same, from raw (e.g. pinned) host pointers."""
        self._n = n_pairs
        self._check(self._L.sfc_batch_upload(self._h, arena_ptr, arena_len, pairs_ptr, n_pairs))

    def run(self, pen: Penalties, want_cigar: bool):
        self._want = bool(want_cigar)
        self._check(self._L.sfc_batch_run(self._h, C.byref(pen), int(want_cigar)))

    def download(self, ops_cap: Optional[int] = None, out=None):
        n = self._n
        if out is None:
            scores = np.empty(n, np.int32)
            status = np.empty(n, np.int32)
        else:
            scores, status = out
        if not self._want:
            self._check(self._L.sfc_batch_download(self._h, scores.ctypes.data, status.ctypes.data, None, 0, None))
            return scores, status, None, None
        ops_off = np.zeros(n + 1, np.uint64)
        cap = ops_cap or max(1 << 16, 64 * n)
        while True:
            ops = np.empty(cap, np.uint32)
            rc = self._L.sfc_batch_download(self._h, scores.ctypes.data, status.ctypes.data, ops.ctypes.data, cap,
                                            ops_off.ctypes.data)
            if rc == -4:
                cap = int(ops_off[n]) + 16
                continue
            self._check(rc)
            return scores, status, ops[:int(ops_off[n])], ops_off

    def stats(self) -> dict:
        st = BatchStats()
        self._check(self._L.sfc_batch_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in BatchStats._fields_}

    # ---- one-shot ----
    def align_batch(self, pen: Penalties, arena: np.ndarray, pairs: np.ndarray, want_cigar: bool):
        self.upload(arena, pairs)
        self.run(pen, want_cigar)
        return self.download()

    def bio_align_batch(self, scoring, arena: np.ndarray, pairs: np.ndarray, want_ops: bool = True):
        """sfc_bio_align_batch: rust-bio style affine alignment with clipping (K4).  pairs: PAIR_DTYPE rows whose
        pattern_* = x and text_* = y.  Returns scores, status, bounds[n,4] (xstart, xend, ystart, yend), ops, ops_off."""
        assert arena.dtype == np.uint8 and arena.flags.c_contiguous
        assert pairs.dtype == PAIR_DTYPE and pairs.flags.c_contiguous
        n = len(pairs)
        scores = np.zeros(n, np.int32)
        status = np.zeros(n, np.int32)
        bounds = np.zeros((n, 4), np.int32)
        if want_ops:
            cap = int(pairs["pattern_len"].astype(np.int64).sum() + pairs["text_len"].astype(np.int64).sum()) + 4 * n
            ops = np.zeros(cap, np.uint32)
            ops_off = np.zeros(n + 1, np.uint64)
            self._check(self._L.sfc_bio_align_batch(self._h, C.byref(scoring), arena.ctypes.data, arena.size, pairs.ctypes.data, n,
                                                    scores.ctypes.data, status.ctypes.data, bounds.ctypes.data, ops.ctypes.data,
                                                    cap, ops_off.ctypes.data))
            return scores, status, bounds, ops, ops_off
        self._check(self._L.sfc_bio_align_batch(self._h, C.byref(scoring), arena.ctypes.data, arena.size, pairs.ctypes.data, n,
                                                scores.ctypes.data, status.ctypes.data, bounds.ctypes.data, None, 0, None))
        return scores, status, bounds, None, None

    def bio_score_batch(self, scoring, arena: np.ndarray, pairs: np.ndarray):
        """sfc_bio_score_batch: the same aligner, score only (K5).  Returns scores, status."""
        assert arena.dtype == np.uint8 and arena.flags.c_contiguous
        assert pairs.dtype == PAIR_DTYPE and pairs.flags.c_contiguous
        n = len(pairs)
        scores = np.zeros(n, np.int32)
        status = np.zeros(n, np.int32)
        self._check(self._L.sfc_bio_score_batch(self._h, C.byref(scoring), arena.ctypes.data, arena.size, pairs.ctypes.data, n,
                                                scores.ctypes.data, status.ctypes.data))
        return scores, status

    def measure_int_peak(self):
        g, m = C.c_double(), C.c_double()
        self._check(self._L.sfc_measure_int_peak(self._h, C.byref(g), C.byref(m)))
        return g.value, m.value


def expand_ops(ops_rle: np.ndarray) -> bytes:
    L = _lib.lib()
    ops_rle = np.ascontiguousarray(ops_rle, np.uint32)
    total = int((ops_rle >> 2).sum())
    buf = np.empty(max(total, 1), np.uint8)
    n = L.sfc_expand_ops(ops_rle.ctypes.data, len(ops_rle), buf.ctypes.data, buf.size)
    if n < 0:
        raise SfcError(int(n), "sfc_expand_ops")
    return buf[:n].tobytes()


def translate_rle(ops_rle: np.ndarray):
    """== process_wfa2_cigar_string on run-length ops. Returns (ref_offset, u32 cigar array, has_match)."""
    L = _lib.lib()
    ops_rle = np.ascontiguousarray(ops_rle, np.uint32)
    cig = np.empty(len(ops_rle) + 1, np.uint32)
    roff, n, hm = C.c_int64(), C.c_size_t(), C.c_int()
    rc = L.sfc_translate_rle(ops_rle.ctypes.data, len(ops_rle), C.byref(roff), cig.ctypes.data, cig.size, C.byref(n),
                             C.byref(hm))
    if rc != 0:
        raise SfcError(rc, "sfc_translate_rle")
    return roff.value, cig[:n.value].copy(), bool(hm.value)


def translate_cigar(wfa_ops: bytes):
    """== process_wfa2_cigar_string (src/wfa2_utils.rs:13-95) on WFA2's byte string."""
    L = _lib.lib()
    buf = np.frombuffer(wfa_ops, np.uint8) if wfa_ops else np.zeros(1, np.uint8)
    cig = np.empty(len(wfa_ops) + 1, np.uint32)
    roff, n, hm = C.c_int64(), C.c_size_t(), C.c_int()
    rc = L.sfc_translate_cigar(buf.ctypes.data, len(wfa_ops), C.byref(roff), cig.ctypes.data, cig.size, C.byref(n),
                               C.byref(hm))
    if rc != 0:
        raise SfcError(rc, "sfc_translate_cigar")
    return roff.value, cig[:n.value].copy(), bool(hm.value)


class MultiGpuContext:
    """One process driving several GPUs (the reference's deployment shape): sfc_multi_* of the C ABI."""

    def __init__(self, devices):
        L = _lib.lib()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = L.sfc_multi_create(devs, len(devices), C.byref(h))
        if rc != 0:
            raise SfcError(rc, "sfc_multi_create failed (B200 devices required; there is no CPU fallback)")
        self._h, self._L, self.devices = h, L, list(devices)

    def close(self):
        if getattr(self, "_h", None):
            self._L.sfc_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align_batch(self, pen: Penalties, arena: np.ndarray, pairs: np.ndarray, want_cigar: bool, group_of_pair=None):
        n = len(pairs)
        scores = np.empty(n, np.int32)
        status = np.empty(n, np.int32)
        ops_off = np.zeros(n + 1, np.uint64)
        grp = None if group_of_pair is None else np.ascontiguousarray(group_of_pair, np.uint32)
        cap = max(1 << 16, 64 * n) if want_cigar else 0
        while True:
            ops = np.empty(max(cap, 1), np.uint32)
            rc = self._L.sfc_multi_align_batch(self._h, C.byref(pen), arena.ctypes.data, arena.size, pairs.ctypes.data, n,
                                               None if grp is None else grp.ctypes.data, int(want_cigar), scores.ctypes.data,
                                               status.ctypes.data, ops.ctypes.data if want_cigar else None, cap,
                                               ops_off.ctypes.data if want_cigar else None)
            if rc == -4 and want_cigar:
                cap = int(ops_off[n]) + 16
                continue
            if rc != 0:
                raise SfcError(rc, (self._L.sfc_multi_last_error(self._h) or b"").decode())
            break
        return scores, status, (ops[:int(ops_off[n])] if want_cigar else None), (ops_off if want_cigar else None)

    def last_kernel_ms(self):
        out = np.zeros(len(self.devices), np.float64)
        self._L.sfc_multi_last_kernel_ms(self._h, out.ctypes.data)
        return out
