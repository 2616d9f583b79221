"""Overlay of the reference's modules/dynamic_modules/RouterTriple.py."""
from dynamicvectorquantization_b200._fallthrough import make_getattr
from dynamicvectorquantization_b200.nn.router import TripleGrainFeatureRouter  # noqa: F401

__getattr__ = make_getattr(__name__, __file__)
