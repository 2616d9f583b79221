"""Overlay of the reference's modules/dynamic_modules/EncoderTriple.py."""
from dynamicvectorquantization_b200.nn.encoder import TripleGrainEncoder  # noqa: F401
