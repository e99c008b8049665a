"""ctypes loader for libsawfish_cuda.so (the C ABI in include/sawfish_cuda.h).

There is no fallback: if the shared library is missing or no B200 is visible, importing the
product API raises.  The oracle under oracle/ is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SFC_SO_PATH") or os.path.join(_PKG, "libsawfish_cuda.so")
CSRC = os.path.join(_PKG, "csrc")

SFC_OK = 0
ERRORS = {-1: "SFC_ERR_INVALID", -2: "SFC_ERR_NO_DEVICE", -3: "SFC_ERR_CUDA", -4: "SFC_ERR_CAPACITY",
          -5: "SFC_ERR_UNSUPPORTED", -6: "SFC_ERR_OOM", -9: "SFC_ERR_INTERNAL"}


class SfcError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Penalties(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("metric", "match", "mismatch", "gap_open1", "gap_ext1", "gap_open2", "gap_ext2")]


class BioScoring(C.Structure):
    """sfc_bio_scoring: rust-bio Scoring (gap_open, gap_extend, match, mismatch) + the four clip penalties."""
    _fields_ = [(n, C.c_int32) for n in ("gap_open", "gap_extend", "match", "mismatch", "xclip_prefix", "xclip_suffix",
                                         "yclip_prefix", "yclip_suffix")]


class BatchStats(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("pack_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("cells", C.c_uint64), ("launches", C.c_uint32),
                ("retries", C.c_uint32), ("n_score_kernel", C.c_uint32), ("n_align_kernel", C.c_uint32)]


# Every symbol include/sawfish_cuda.h declares; tests check the .so exports all of them.
EXPORTS = [
    "sfc_abi_version", "sfc_device_count", "sfc_create", "sfc_destroy", "sfc_last_error", "sfc_set_workspace_limit",
    "sfc_align_batch", "sfc_batch_upload", "sfc_batch_run", "sfc_batch_download", "sfc_batch_get_stats",
    "sfc_expand_ops", "sfc_translate_cigar", "sfc_translate_rle",
    "sfc_aligner_new_consensus", "sfc_aligner_new_large_indel", "sfc_aligner_free", "sfc_aligner_set_text",
    "sfc_aligner_align", "sfc_aligner_matching", "sfc_measure_int_peak", "sfc_estimate_pair_work", "sfc_shard_plan", "sfc_multi_create", "sfc_multi_destroy",
    "sfc_multi_device_count", "sfc_multi_last_error", "sfc_multi_align_batch", "sfc_multi_last_kernel_ms",
    "sfc_bio_align_batch", "sfc_bio_score_batch",
]


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_PKG), "include", "sawfish_cuda.h"))
    if force or not os.path.exists(SO_PATH) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs):
        subprocess.run(["make", "-C", CSRC, "-s"], check=True)
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, u8p, i32p, u32p, u64p, i64p, szp = (C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_int64), C.POINTER(C.c_size_t))
    L.sfc_abi_version.restype = C.c_int
    L.sfc_device_count.restype = C.c_int
    L.sfc_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.sfc_destroy.argtypes = [vp]
    L.sfc_destroy.restype = None
    L.sfc_last_error.argtypes = [vp]
    L.sfc_last_error.restype = C.c_char_p
    L.sfc_set_workspace_limit.argtypes = [vp, C.c_size_t]
    L.sfc_align_batch.argtypes = [vp, C.POINTER(Penalties), u8p, C.c_size_t, vp, C.c_size_t, C.c_int, i32p, i32p, u32p,
                                  C.c_size_t, u64p]
    L.sfc_batch_upload.argtypes = [vp, u8p, C.c_size_t, vp, C.c_size_t]
    L.sfc_batch_run.argtypes = [vp, C.POINTER(Penalties), C.c_int]
    L.sfc_batch_download.argtypes = [vp, i32p, i32p, u32p, C.c_size_t, u64p]
    L.sfc_batch_get_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.sfc_expand_ops.argtypes = [u32p, C.c_size_t, u8p, C.c_size_t]
    L.sfc_expand_ops.restype = C.c_int64
    L.sfc_translate_cigar.argtypes = [u8p, C.c_size_t, i64p, u32p, C.c_size_t, szp, C.POINTER(C.c_int)]
    L.sfc_translate_rle.argtypes = [u32p, C.c_size_t, i64p, u32p, C.c_size_t, szp, C.POINTER(C.c_int)]
    L.sfc_aligner_new_consensus.argtypes = [vp, C.POINTER(vp)]
    L.sfc_aligner_new_large_indel.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.sfc_aligner_free.argtypes = [vp]
    L.sfc_aligner_free.restype = None
    L.sfc_aligner_set_text.argtypes = [vp, u8p, C.c_size_t]
    L.sfc_aligner_align.argtypes = [vp, u8p, C.c_size_t, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int), i64p, u32p, C.c_size_t, szp]
    L.sfc_aligner_matching.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, szp]
    L.sfc_measure_int_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sfc_estimate_pair_work.argtypes = [C.POINTER(Penalties), vp, C.c_size_t, vp]
    L.sfc_shard_plan.argtypes = [vp, C.c_size_t, C.c_int, vp]
    L.sfc_multi_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.sfc_multi_destroy.argtypes = [vp]
    L.sfc_multi_destroy.restype = None
    L.sfc_multi_device_count.argtypes = [vp]
    L.sfc_multi_last_error.argtypes = [vp]
    L.sfc_multi_last_error.restype = C.c_char_p
    L.sfc_multi_align_batch.argtypes = [vp, C.POINTER(Penalties), u8p, C.c_size_t, vp, C.c_size_t, vp, C.c_int, i32p, i32p, u32p,
                                        C.c_size_t, u64p]
    L.sfc_multi_last_kernel_ms.argtypes = [vp, vp]
    L.sfc_bio_align_batch.argtypes = [vp, C.POINTER(BioScoring), u8p, C.c_size_t, vp, C.c_size_t, i32p, i32p, i32p, u32p,
                                      C.c_size_t, u64p]
    L.sfc_bio_score_batch.argtypes = [vp, C.POINTER(BioScoring), u8p, C.c_size_t, vp, C.c_size_t, i32p, i32p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("sfc_abi_version", "sfc_device_count"):
            fn.restype = C.c_int
    _lib = L
    return L
