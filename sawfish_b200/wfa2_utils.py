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

    def matching(self, text: bytes = b"", pattern: bytes = b"") -> Tuple[str, str, str]:
        """Debug view of the last alignment (src/wfa2_utils.rs:253-255): gapped pattern, '|' under matches, gapped text,
        in the order the sequences were aligned (both reversed).  The arguments are accepted for signature parity; like
        WFA2's `matching` the lines come from the aligner's stored cigar."""
        L = _lib.lib()
        n = C.c_size_t()
        rc = L.sfc_aligner_matching(self._h, None, None, None, 0, C.byref(n))
        if n.value == 0:
            if rc not in (0, -4):
                raise SfcError(rc, "matching: no alignment yet")
            return "", "", ""
        bufs = [C.create_string_buffer(n.value) for _ in range(3)]
        rc = L.sfc_aligner_matching(self._h, bufs[0], bufs[1], bufs[2], n.value, C.byref(n))
        if rc != 0:
            raise SfcError(rc, "matching")
        return tuple(b.raw[:n.value].decode() for b in bufs)


def get_pairwise_alignment_mismatch_line(lines, start: int, end: int) -> str:
    """src/utils.rs:139-166: '|' match, 'X' mismatch, ' ' gap between the first two lines."""
    out = []
    for pos in range(start, end):
        a, b = lines[0][pos:pos + 1], lines[1][pos:pos + 1]
        out.append(" " if a == b"-" or b == b"-" else ("|" if a == b else "X"))
    return "".join(out)


def pairwise_alignment_printer(width: int, lines, file=None) -> None:
    """src/utils.rs:174-213: wraps a pairwise alignment over rows of `width` columns with a ruler and a mismatch line."""
    import sys
    file = file or sys.stderr
    assert len(lines) > 1
    n = len(lines[0])
    assert len(lines[1]) == n
    ruler = "".join("." if (i + 1) % 10 == 0 else " " for i in range(width))
    for row in range((n + width - 1) // width):
        start, end = width * row, min(width * row + width, n)
        print(ruler, file=file)
        print(bytes(lines[0][start:end]).decode(), file=file)
        print(get_pairwise_alignment_mismatch_line(lines, start, end), file=file)
        print(bytes(lines[1][start:end]).decode(), file=file)
        print(file=file)


def print_pairwise_alignment(aligner: "PairwiseAligner", pattern: bytes, file=None) -> None:
    """src/wfa2_utils.rs:258-273: text over pattern, in forward orientation, 150 columns per row."""
    l1, _, l3 = aligner.matching(aligner.text(), bytes(pattern)[::-1])
    pairwise_alignment_printer(150, [l3.encode()[::-1], l1.encode()[::-1]], file=file)

