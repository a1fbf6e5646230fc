"""dfmir_b200 — B200-native translation + registration hot path of heyblackC/DFMIR.

Host side mirrors the reference's Python surface; compute runs in
libdfmir_b200.so (hand-written sm_100a CUDA behind the C ABI of include/dfmir_b200.h).
"""
from . import _lib  # noqa: F401
from .layers import SpatialTransformer, VecInt, ResizeTransform  # noqa: F401
from .losses import NCC_Loss, Grad_Loss, smooothing_loss, calculate_L1_loss  # noqa: F401

from . import functional, networks, vxm, patchnce, registration_model, fused, vxm_trainer  # noqa: F401,E402
from .fused import integrate_warp_loss  # noqa: F401,E402
from .networks import define_G, define_F, ResnetGenerator, PatchSampleF  # noqa: F401,E402
from .vxm import VxmDense, Unet  # noqa: F401,E402
from .patchnce import PatchNCELoss  # noqa: F401,E402
from .registration_model import REGISTRATIONModel, RegistrationModel  # noqa: F401,E402
from .vxm_trainer import VxmRegistrationTrainer  # noqa: F401,E402

__all__ = ["define_G", "define_F", "ResnetGenerator", "PatchSampleF", "VxmDense", "Unet", "PatchNCELoss",
           "REGISTRATIONModel", "RegistrationModel", "VxmRegistrationTrainer",
           "SpatialTransformer", "VecInt", "ResizeTransform", "NCC_Loss", "Grad_Loss",
           "smooothing_loss", "calculate_L1_loss", "integrate_warp_loss"]
