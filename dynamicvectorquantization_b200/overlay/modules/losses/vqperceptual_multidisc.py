"""Overlay of the reference's modules/losses/vqperceptual_multidisc.py."""
from dynamicvectorquantization_b200._fallthrough import make_getattr
from dynamicvectorquantization_b200.nn.losses import (DummyLoss, VQLPIPSWithDiscriminator, adopt_weight,  # noqa: F401
                                                      bce_discr_loss, bce_gen_loss, hinge_d_loss, hinge_g_loss,
                                                      vanilla_d_loss)

__getattr__ = make_getattr(__name__, __file__)
