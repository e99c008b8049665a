"""Mirror of the reference's src/wfa2_utils.rs call surface on top of the CUDA library.

Same names, argument meaning and error behaviour as wfa2_utils::PairwiseAligner (src/wfa2_utils.rs:143-256):
asserts on empty text/pattern, (score, Option<SimpleAlignment>) results, sequences reversed before alignment
and the CIGAR re-labelled afterwards.  Each call is an n=1 batch through the C ABI (sfc_aligner_*); batching
callers use sawfish_b200.batch.AlignContext directly.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import SfcError
from .batch import AlignContext, translate_cigar
from .simple_alignment import SimpleAlignment, cigar_from_u32

_default_ctx: Optional[AlignContext] = None


def default_context() -> AlignContext:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = AlignContext(0)
    return _default_ctx


def process_wfa2_cigar_string(wfa2_cigar: bytes) -> Optional[SimpleAlignment]:
    """src/wfa2_utils.rs:13-95."""
    roff, cig, has_match = translate_cigar(bytes(wfa2_cigar))
    if not has_match:
        return None
    return SimpleAlignment(roff, cigar_from_u32(cig))


class PairwiseAligner:
    def __init__(self, handle, ctx: AlignContext):
        self._h = handle
        self._ctx = ctx
        self._text = b""

    @classmethod
    def _new(cls, ctor: str, ctx: Optional[AlignContext], *args) -> "PairwiseAligner":
        ctx = ctx or default_context()
        L = _lib.lib()
        h = C.c_void_p()
        rc = getattr(L, ctor)(ctx._h, *args, C.byref(h))
        if rc != 0:
            raise SfcError(rc, ctor)
        return cls(h, ctx)

    @classmethod
    def new_large_indel_aligner(cls, is_long_contig: bool, ctx: Optional[AlignContext] = None) -> "PairwiseAligner":
        return cls._new("sfc_aligner_new_large_indel", ctx, int(is_long_contig))

    @classmethod
    def new_consensus_aligner(cls, ctx: Optional[AlignContext] = None) -> "PairwiseAligner":
        return cls._new("sfc_aligner_new_consensus", ctx)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().sfc_aligner_free(self._h)
                self._h = None
        except Exception:
            pass

    def set_text(self, text: bytes):
        assert len(text) > 0  # assert!(!text.is_empty())  src/wfa2_utils.rs:199
        self._text = bytes(text)
        buf = np.frombuffer(self._text, np.uint8)
        rc = _lib.lib().sfc_aligner_set_text(self._h, buf.ctypes.data, len(self._text))
        if rc != 0:
            raise SfcError(rc, "set_text")

    def text(self) -> bytes:
        """The stored (reversed) text, as the reference returns it (src/wfa2_utils.rs:205-207)."""
        return self._text[::-1]

    def _align(self, pattern: bytes, mode: int, tcf: float = 0.0, pcf: float = 0.0) -> Tuple[int, Optional[SimpleAlignment]]:
        assert len(self._text) > 0   # src/wfa2_utils.rs:210
        assert len(pattern) > 0      # src/wfa2_utils.rs:211
        pattern = bytes(pattern)
        buf = np.frombuffer(pattern, np.uint8)
        cap = len(pattern) + len(self._text) + 8
        cig = np.empty(cap, np.uint32)
        score, hm, roff, n = C.c_int32(), C.c_int(), C.c_int64(), C.c_size_t()
        rc = _lib.lib().sfc_aligner_align(self._h, buf.ctypes.data, len(pattern), mode, tcf, pcf, C.byref(score),
                                          C.byref(hm), C.byref(roff), cig.ctypes.data, cap, C.byref(n))
        if rc != 0:
            # the reference asserts StatusAlgCompleted (src/wfa2_utils.rs:117) and aborts; raise instead of returning junk
            raise SfcError(rc, (_lib.lib().sfc_last_error(self._ctx._h) or b"").decode())
        al = SimpleAlignment(roff.value, cigar_from_u32(cig[:n.value])) if hm.value else None
        return score.value, al

    def align(self, pattern: bytes):
        return self._align(pattern, 0)

    def align_clip_frac(self, pattern: bytes, text_clip_frac: float, pattern_clip_frac: float):
        return self._align(pattern, 1, text_clip_frac, pattern_clip_frac)

    def align_end_to_end(self, pattern: bytes):
        return self._align(pattern, 2)
