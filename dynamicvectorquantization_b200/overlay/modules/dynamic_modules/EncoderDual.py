"""Overlay of the reference's modules/dynamic_modules/EncoderDual.py."""
from dynamicvectorquantization_b200.nn.encoder import DualGrainEncoder  # noqa: F401
