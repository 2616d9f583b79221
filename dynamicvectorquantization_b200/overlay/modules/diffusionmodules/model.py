"""Overlay of the reference's modules/diffusionmodules/model.py (hot-path classes only)."""
from dynamicvectorquantization_b200._fallthrough import make_getattr
from dynamicvectorquantization_b200.nn.blocks import (AttnBlock, Downsample, Normalize, ResnetBlock,  # noqa: F401
                                                      Upsample, nonlinearity)

__getattr__ = make_getattr(__name__, __file__)
