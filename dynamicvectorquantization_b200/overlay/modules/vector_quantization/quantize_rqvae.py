"""Overlay of the reference's modules/vector_quantization/quantize_rqvae.py."""
from dynamicvectorquantization_b200.nn.quantize import VQEmbedding  # noqa: F401
from dynamicvectorquantization_b200.nn.quantize_family import RQBottleneck  # noqa: F401
