"""rcot_b200 -- B200-native (sm_100a) kernels for RCOT's training hot path.

Host side is Python/PyTorch; the compute runs in ``librcot_b200.so`` (hand-written CUDA, C-ABI
declared in ``include/rcot_b200.h``).  No CPU fallback exists by design.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
