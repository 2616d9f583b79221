"""Overlay of the reference's models/stage1_dynamic/dqvae_triple_feat.py."""
from dynamicvectorquantization_b200.nn.model import TripleGrainVQModel  # noqa: F401
