"""Overlay of the reference's modules/vector_quantization/quantize2_mask.py."""
from dynamicvectorquantization_b200.nn.quantize import VQEmbedding, VectorQuantize2  # noqa: F401
