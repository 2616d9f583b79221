"""Overlay of the reference's modules/dynamic_modules/permuter.py."""
from dynamicvectorquantization_b200.nn.permuter import DualGrainSeperatePermuter  # noqa: F401
