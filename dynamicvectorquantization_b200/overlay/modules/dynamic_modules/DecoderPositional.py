"""Overlay of the reference's modules/dynamic_modules/DecoderPositional.py."""
from dynamicvectorquantization_b200.nn.decoder import Decoder, PositionEmbedding2DLearned  # noqa: F401
