"""rcot_b200 -- B200-native (sm_100a) kernels for RCOT's training hot path.

Host side is Python/PyTorch; the compute runs in ``librcot_b200.so`` (hand-written CUDA, C-ABI
declared in ``include/rcot_b200.h``).  No CPU fallback exists by design.
"""
from . import _lib  # noqa: F401


def set_hidden_dtype(name):
    """'fp32' (default) or 'bf16': storage type of the Restormer blocks' HIDDEN tensors (pre, qkv, u, g and their
    gradients) on the levels with C <= 96.  Block inputs/outputs, weights, LayerNorm statistics, every accumulation and
    every weight gradient stay fp32; products of a bf16-stored operand need no hi/lo split (it is exact in bf16).
    Stated tolerance of the bf16 mode: network output rtol 2e-2 / atol 2e-3, losses 2e-2 (tests/test_bf16_mode.py)."""
    import torch

    from . import ops
    if name not in ("fp32", "bf16"):
        raise ValueError(f"hidden dtype must be 'fp32' or 'bf16', got {name!r}")
    ops.HIDDEN_DTYPE = torch.bfloat16 if name == "bf16" else torch.float32


__all__ = ["_lib", "set_hidden_dtype"]
