"""PatchGAN discriminator of the stage-1 training step.

Mirror of ``modules/discriminator/model.py`` (reference): ``weights_init`` (:8-14) and
``NLayerDiscriminator`` (:17-67) with the same constructor arguments and ``state_dict`` keys
(``main.{0,2,5,8,11}.weight``, BatchNorm at ``main.{3,6,9}``).

Round-1 status: the network is ~6 GFLOP per 256x256 image (1.5 % of the
autoencoder's forward) and is made of 4x4 stride-2 convolutions, training-mode BatchNorm and LeakyReLU,
none of which the hand-written kernels cover yet; it runs on PyTorch's CUDA ops (TF32 allowed, as
torch's cuDNN default).  The perceptual term - the FLOP-heavy part of the loss - is on the tensor-core
path (``nn/lpips.py``).
"""
import functools

import torch.nn as nn


def weights_init(m):
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm") != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class NLayerDiscriminator(nn.Module):
    """PatchGAN discriminator as in Pix2Pix (discriminator/model.py:17-67)."""

    def __init__(self, input_nc=3, ndf=64, n_layers=3, use_actnorm=False):
        super().__init__()
        if not use_actnorm:
            norm_layer = nn.BatchNorm2d
        else:
            try:
                from utils.utils import ActNorm          # reference tree (falls through the overlay)
            except Exception as e:
                raise NotImplementedError("use_actnorm=True needs the reference's utils.utils.ActNorm on sys.path") from e
            norm_layer = ActNorm
        if type(norm_layer) == functools.partial:
            use_bias = norm_layer.func != nn.BatchNorm2d
        else:
            use_bias = norm_layer != nn.BatchNorm2d
        kw, padw = 4, 1
        sequence = [nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(0.2, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** n, 8)
            sequence += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw,
                                   bias=use_bias),
                         norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        sequence += [nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias),
                     norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]
        sequence += [nn.Conv2d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw)]
        self.main = nn.Sequential(*sequence)

    def forward(self, input):
        return self.main(input)
