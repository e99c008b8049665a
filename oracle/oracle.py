"""ctypes binding of the CPU oracle (oracle/libwfa_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (sawfish_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwfa_oracle.so")

MODE_LITERAL = 0
MODE_COMPACT = 1


class Penalties(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("metric", "match", "mismatch", "gap_open1", "gap_ext1", "gap_open2", "gap_ext2")]


class Stats(C.Structure):
    _fields_ = [("cells", C.c_int64), ("ext_words", C.c_int64), ("bt_bytes", C.c_int64), ("wf_score", C.c_int32),
                ("end_k", C.c_int32), ("end_off", C.c_int32), ("k0", C.c_int32), ("n_ops", C.c_int32)]


PAIR_DTYPE = np.dtype([("pattern_off", "<u8"), ("text_off", "<u8"), ("pattern_len", "<u4"), ("text_len", "<u4"),
                       ("pattern_begin_free", "<i4"), ("pattern_end_free", "<i4"), ("text_begin_free", "<i4"),
                       ("text_end_free", "<i4"), ("flags", "<u4")], align=True)
STATS_DTYPE = np.dtype([("cells", "<i8"), ("ext_words", "<i8"), ("bt_bytes", "<i8"), ("wf_score", "<i4"),
                        ("end_k", "<i4"), ("end_off", "<i4"), ("k0", "<i4"), ("n_ops", "<i4")], align=True)

# reference src/wfa2_utils.rs:183-189 and :159-168
CONSENSUS = Penalties(0, -1, 3, 0, 2, 0, 0)
LARGE_INDEL = Penalties(1, -1, 3, 1, 2, 60, 0)


BIO_MIN_SCORE = -858993459


class BioScoring(C.Structure):
    """rust-bio Scoring::from_scores(gap_open, gap_extend, match, mismatch) + clip penalties (MIN_SCORE = disabled)."""
    _fields_ = [(n, C.c_int32) for n in ("gap_open", "gap_extend", "match", "mismatch", "xclip_prefix", "xclip_suffix",
                                         "yclip_prefix", "yclip_suffix")]


class BioResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("score", "xstart", "xend", "ystart", "yend")]


def bio_scoring(gap_open, gap_extend, match, mismatch, xclip_prefix=BIO_MIN_SCORE, xclip_suffix=BIO_MIN_SCORE,
                yclip_prefix=BIO_MIN_SCORE, yclip_suffix=BIO_MIN_SCORE) -> BioScoring:
    return BioScoring(gap_open, gap_extend, match, mismatch, xclip_prefix, xclip_suffix, yclip_prefix, yclip_suffix)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("wfa_oracle.c", "dp_verify.c", "bio_oracle.c", "wfa_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p = C.POINTER(C.c_uint8)
        L.ora_wfa_align.argtypes = [C.POINTER(Penalties), u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.POINTER(C.c_int32), u8p, C.c_int, C.POINTER(Stats)]
        L.ora_wfa_align.restype = C.c_int
        L.ora_translate_cigar.argtypes = [u8p, C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.c_size_t,
                                          C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.ora_translate_cigar.restype = C.c_int
        L.ora_sawfish_align.argtypes = [C.POINTER(Penalties), u8p, C.c_int, u8p, C.c_int, C.c_double, C.c_double, C.c_int,
                                        C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint32),
                                        C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(Stats)]
        L.ora_sawfish_align.restype = C.c_int
        L.ora_align_batch.argtypes = [C.POINTER(Penalties), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ora_align_batch.restype = C.c_int
        L.ora_max_threads.restype = C.c_int
        L.dp_min_penalty.argtypes = [C.c_int] * 6 + [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.dp_min_penalty.restype = C.c_int
        L.bio_custom_align.argtypes = [C.POINTER(BioScoring), u8p, C.c_int, u8p, C.c_int, C.POINTER(BioResult),
                                       C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_int)]
        L.bio_custom_align.restype = C.c_int
        _lib = L
    return _lib


def _buf(b: bytes):
    return (C.c_uint8 * len(b)).from_buffer_copy(b)


CIGAR_CHARS = "MIDNSHP=X"


def cigar_to_string(cigar) -> str:
    return "".join(f"{int(c) >> 4}{CIGAR_CHARS[int(c) & 0xF]}" for c in cigar)


@dataclass
class WfaResult:
    score: int
    ops: bytes
    stats: dict


def wfa_align(pen: Penalties, pattern: bytes, text: bytes, pbf: int, pef: int, tbf: int, tef: int,
              mode: int = MODE_LITERAL) -> WfaResult:
    """Core WFA in the aligner's own coordinates (no reversal)."""
    L = lib()
    score = C.c_int32()
    cap = len(pattern) + len(text) + 8
    ops = (C.c_uint8 * cap)()
    st = Stats()
    rc = L.ora_wfa_align(C.byref(pen), _buf(pattern), len(pattern), _buf(text), len(text), pbf, pef, tbf, tef, mode, 1,
                         C.byref(score), ops, cap, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"ora_wfa_align failed rc={rc}")
    return WfaResult(score.value, bytes(ops[:st.n_ops]), {f: getattr(st, f) for f, _ in Stats._fields_})


def translate_cigar(ops: bytes):
    """== process_wfa2_cigar_string. Returns (ref_offset, [u32 cigar], has_match)."""
    L = lib()
    roff = C.c_int64()
    n = C.c_size_t()
    hm = C.c_int()
    cap = len(ops) + 1
    cig = (C.c_uint32 * cap)()
    rc = L.ora_translate_cigar(_buf(ops) if ops else (C.c_uint8 * 1)(), len(ops), C.byref(roff), cig, cap, C.byref(n),
                               C.byref(hm))
    if rc != 0:
        raise RuntimeError(f"ora_translate_cigar rc={rc}")
    return roff.value, list(cig[:n.value]), bool(hm.value)


def sawfish_align(pen: Penalties, text: bytes, pattern: bytes, text_clip_frac: float = -1.0,
                  pattern_clip_frac: float = -1.0, end_to_end: bool = False, mode: int = MODE_LITERAL):
    """== wfa2_utils::PairwiseAligner::{set_text + align}. Returns (score, None | (ref_offset, cigar u32 list), stats)."""
    L = lib()
    score = C.c_int32()
    roff = C.c_int64()
    n = C.c_size_t()
    hm = C.c_int()
    cap = len(pattern) + len(text) + 8
    cig = (C.c_uint32 * cap)()
    st = Stats()
    rc = L.ora_sawfish_align(C.byref(pen), _buf(text), len(text), _buf(pattern), len(pattern), text_clip_frac,
                             pattern_clip_frac, int(end_to_end), mode, C.byref(score), C.byref(roff), cig, cap,
                             C.byref(n), C.byref(hm), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"ora_sawfish_align rc={rc}")
    al = (roff.value, list(cig[:n.value])) if hm.value else None
    return score.value, al, {f: getattr(st, f) for f, _ in Stats._fields_}


def align_batch(pen: Penalties, arena: np.ndarray, pairs: np.ndarray, want_ops: bool = True,
                mode: int = MODE_COMPACT, n_threads: int = 0):
    """Batch driver (CPU baseline). Returns scores, status, stats, ops(bytes array), ops_off."""
    L = lib()
    assert pairs.dtype == PAIR_DTYPE and arena.dtype == np.uint8
    n = len(pairs)
    scores = np.zeros(n, np.int32)
    status = np.zeros(n, np.int32)
    stats = np.zeros(n, STATS_DTYPE)
    ops_off = np.zeros(n + 1, np.uint64)
    cap = int(pairs["pattern_len"].astype(np.int64).sum() + pairs["text_len"].astype(np.int64).sum())
    ops = np.zeros(cap if want_ops else 1, np.uint8)
    rc = L.ora_align_batch(C.byref(pen), arena.ctypes.data, pairs.ctypes.data, n, int(want_ops), mode, n_threads,
                           scores.ctypes.data, status.ctypes.data, stats.ctypes.data,
                           ops.ctypes.data if want_ops else None, cap, ops_off.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"ora_align_batch rc={rc}")
    return scores, status, stats, ops, ops_off


def adjusted(pen: Penalties):
    m = pen.match
    if m == 0:
        x, o1, e1, o2, e2 = pen.mismatch, pen.gap_open1, pen.gap_ext1, pen.gap_open2, pen.gap_ext2
    else:
        x = 2 * pen.mismatch - 2 * m
        o1, e1, o2, e2 = 2 * pen.gap_open1, 2 * pen.gap_ext1 - m, 2 * pen.gap_open2, 2 * pen.gap_ext2 - m
    if pen.metric == 0:
        o1 = 0
    return x, o1, e1, o2, e2


def dp_min_penalty(pen: Penalties, pattern: bytes, text: bytes, pbf: int, pef: int, tbf: int, tef: int) -> int:
    x, o1, e1, o2, e2 = adjusted(pen)
    return lib().dp_min_penalty(pen.metric, x, o1, e1, o2, e2, _buf(pattern), len(pattern), _buf(text), len(text), pbf,
                                pef, tbf, tef)


def max_threads() -> int:
    return lib().ora_max_threads()


BIO_OPS = ("Match", "Subst", "Ins", "Del", "Xclip", "Yclip")


def bio_custom_align(scoring: BioScoring, x: bytes, y: bytes):
    """rust-bio Aligner::custom(x = pattern, y = text) restated (oracle/bio_oracle.c).
    Returns (score, xstart, xend, ystart, yend, [(op name, len)]) with one entry per operation."""
    L = lib()
    cap = 2 * (len(x) + len(y)) + 16
    ops = (C.c_uint32 * cap)()
    n_ops = C.c_int(0)
    res = BioResult()
    rc = L.bio_custom_align(C.byref(scoring), _buf(x), len(x), _buf(y), len(y), C.byref(res), ops, cap, C.byref(n_ops))
    if rc != 0:
        raise RuntimeError(f"bio_custom_align rc={rc}")
    out = [(BIO_OPS[w & 7], w >> 3) for w in ops[:n_ops.value]]
    return res.score, res.xstart, res.xend, res.ystart, res.yend, out
