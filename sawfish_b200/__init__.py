"""sawfish_b200 — B200-native (sm_100a) drop-in for sawfish's WFA pairwise-alignment hot path.

Product code only: the CUDA library (csrc/ -> libsawfish_cuda.so, C ABI in include/sawfish_cuda.h) and
host-side mirrors of the reference interface (wfa2_utils, simple_alignment, batch).  Nothing here imports
oracle/.
"""
from ._lib import SO_PATH, SfcError, build  # noqa: F401

__all__ = ["SO_PATH", "SfcError", "build"]
