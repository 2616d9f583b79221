"""Overlay of the reference's modules/vector_quantization/quantize_vqgan.py: VectorQuantizer2 runs on the
sm_100a search kernel; the other classes of that file (VectorQuantizer, GumbelQuantize, EMAVectorQuantizer)
are not on the path and resolve to the reference's own definitions when its tree is importable."""
from dynamicvectorquantization_b200._fallthrough import make_getattr
from dynamicvectorquantization_b200.nn.quantize_family import VectorQuantizer2  # noqa: F401

__getattr__ = make_getattr(__name__, __file__)
