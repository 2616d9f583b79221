"""Overlay of the reference's modules/losses/lpips.py."""
from dynamicvectorquantization_b200._fallthrough import make_getattr
from dynamicvectorquantization_b200.nn.lpips import (LPIPS, NetLinLayer, ScalingLayer, normalize_tensor,  # noqa: F401
                                                     spatial_average, vgg16)

__getattr__ = make_getattr(__name__, __file__)
