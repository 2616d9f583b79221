"""Overlay of the reference's modules/vector_quantization/quantize2_list.py."""
from dynamicvectorquantization_b200.nn.quantize import VQEmbedding  # noqa: F401
from dynamicvectorquantization_b200.nn.quantize_family import VectorQuantize2List as VectorQuantize2  # noqa: F401
