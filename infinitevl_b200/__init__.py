"""infinitevl_b200 -- B200-native (sm_100a) hybrid-attention hot path of InfiniteVL.

Gated DeltaNet chunk scan + Sliding-Window Attention as hand-written CUDA behind a C ABI
(include/ivl_b200.h), with the reference's operator / cache / mixer interfaces on top.
Importing the package does not need a GPU; calling an operator without the compiled
library or without a CUDA device raises -- there is no CPU fallback.
"""
__version__ = "0.1.0"

from ._lib import IvlError, LIB_PATH, load  # noqa: F401
