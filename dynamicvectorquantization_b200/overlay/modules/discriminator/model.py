"""Overlay of the reference's modules/discriminator/model.py."""
from dynamicvectorquantization_b200._fallthrough import make_getattr
from dynamicvectorquantization_b200.nn.discriminator import NLayerDiscriminator, weights_init  # noqa: F401

__getattr__ = make_getattr(__name__, __file__)
