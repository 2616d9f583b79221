"""Overlay of the reference's models/stage1_dynamic/dqvae_dual_feat.py."""
from dynamicvectorquantization_b200.nn.model import DualGrainVQModel  # noqa: F401
