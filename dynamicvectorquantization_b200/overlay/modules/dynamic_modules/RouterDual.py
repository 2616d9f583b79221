"""Overlay of the reference's modules/dynamic_modules/RouterDual.py."""
from dynamicvectorquantization_b200.nn.router import DualGrainFeatureRouter, DualGrainFixedEntropyRouter  # noqa: F401
