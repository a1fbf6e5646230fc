"""dfmir_b200 — B200-native translation + registration hot path of heyblackC/DFMIR.

Host side mirrors the reference's Python surface; compute runs in
libdfmir_b200.so (hand-written sm_100a CUDA behind the C ABI of include/dfmir_b200.h).
"""
from . import _lib  # noqa: F401
from .layers import SpatialTransformer, VecInt, ResizeTransform  # noqa: F401
from .losses import NCC_Loss, Grad_Loss, smooothing_loss, calculate_L1_loss  # noqa: F401

__all__ = ["SpatialTransformer", "VecInt", "ResizeTransform", "NCC_Loss", "Grad_Loss",
           "smooothing_loss", "calculate_L1_loss"]
