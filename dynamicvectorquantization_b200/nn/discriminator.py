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
import torch.nn as nn


def weights_init(m):
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm") != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class NLayerDiscriminator(nn.Module):
    """PatchGAN discriminator as in Pix2Pix (discriminator/model.py:17-67): a stem convolution, `n_layers - 1`
    stride-2 stages and one stride-1 stage of conv4x4 -> norm -> LeakyReLU(0.2) with widths ndf * min(2^n, 8), and a
    one-channel conv4x4 head.  ``main`` holds the same modules at the same indices as the reference's Sequential."""

    def __init__(self, input_nc=3, ndf=64, n_layers=3, use_actnorm=False):
        super().__init__()
        if use_actnorm:
            try:
                from utils.utils import ActNorm          # reference tree (falls through the overlay)
            except Exception as e:
                raise NotImplementedError("use_actnorm=True needs the reference's utils.utils.ActNorm on sys.path") from e
            norm_layer = ActNorm
        else:
            norm_layer = nn.BatchNorm2d
        conv_bias = norm_layer is not nn.BatchNorm2d      # BatchNorm brings its own shift (:30-33)
        widths = [ndf * min(2 ** n, 8) for n in range(n_layers + 1)]
        strides = [2] * (n_layers - 1) + [1]

        def conv(cin, cout, stride, bias=True):
            return nn.Conv2d(cin, cout, kernel_size=4, stride=stride, padding=1, bias=bias)

        layers = [conv(input_nc, widths[0], 2), nn.LeakyReLU(0.2, True)]
        for cin, cout, stride in zip(widths[:-1], widths[1:], strides):
            layers += [conv(cin, cout, stride, bias=conv_bias), norm_layer(cout), nn.LeakyReLU(0.2, True)]
        layers.append(conv(widths[-1], 1, 1))             # one prediction per receptive-field patch
        self.main = nn.Sequential(*layers)

    def forward(self, input):
        return self.main(input)
