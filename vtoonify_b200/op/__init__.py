"""Mirror of the reference package ``model/stylegan/op/__init__.py:1-2`` (+ conv2d_gradfix, imported at
model/stylegan/model.py:11)."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d
from . import conv2d_gradfix

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "conv2d_gradfix"]
