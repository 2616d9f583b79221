"""Overlay of the reference's models/stage1_dynamic/dqvae_dual_entropy.py."""
from dynamicvectorquantization_b200.nn.model import DualGrainEntropyVQModel as DualGrainVQModel  # noqa: F401
from dynamicvectorquantization_b200.nn.model import Entropy  # noqa: F401
