"""ResNet / attention / resampling blocks of the DQ-VAE conv stacks on the sm_100a kernels.

Mirror of ``modules/diffusionmodules/model.py:29-192`` of the reference: same class names,
constructor arguments, sub-module names (=> identical ``state_dict`` keys and default init order)
and forward signatures.  ``forward`` takes/returns the reference's NCHW fp32 tensors;
``forward_nhwc`` is the internal NHWC bf16 path the encoder/decoder chain through.
"""
import torch
import torch.nn as nn

from .. import ops


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("dynamicvectorquantization_b200 runs on CUDA (sm_100a) only; "
                           "there is no CPU fallback for the DQ-VAE hot path")


def nonlinearity(x):
    """swish (model.py:29-31); kept for callers that apply it outside a fused GroupNorm."""
    return x * torch.sigmoid(x)


def Normalize(in_channels):
    """model.py:34-35.  The nn.GroupNorm object is a parameter container; the kernels read its
    weight/bias (see ops.gn_swish)."""
    return torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class _Nhwc(nn.Module):
    def forward(self, x, *args, **kwargs):
        _require_cuda(x)
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class Upsample(_Nhwc):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if self.with_conv:
            self.conv = torch.nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward_nhwc(self, x):
        if self.with_conv and ops.FOLD_UPSAMPLE and x.shape[-1] % 64 == 0 and self.conv.out_channels % 64 == 0:
            # nearest x2 + conv3x3 as four 2x2 convolutions of the low-resolution input (2.25x fewer FLOPs)
            return ops.UpsampleConvFn.apply(x, self.conv.weight, self.conv.bias)
        x = ops.Upsample2xFn.apply(x)
        if self.with_conv:
            x = ops.conv2d(x, self.conv)
        return x


class Downsample(_Nhwc):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if self.with_conv:
            # pad (0,1,0,1) + stride 2 is folded into the kernel's TMA coordinates
            self.conv = torch.nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward_nhwc(self, x):
        if not self.with_conv:
            raise NotImplementedError("avg-pool down-sampling is not used by the stage-1 configs")
        return ops.conv2d(x, self.conv, stride=2)


class ResnetBlock(_Nhwc):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = torch.nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if temb_channels > 0:
            self.temb_proj = torch.nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = torch.nn.Dropout(dropout)
        self.conv2 = torch.nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            if self.use_conv_shortcut:
                self.conv_shortcut = torch.nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = torch.nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, temb=None):
        if temb is not None:
            raise NotImplementedError("timestep embeddings are not part of the stage-1 path (temb_ch=0)")
        return super().forward(x)

    def forward_nhwc(self, x):
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("dropout > 0 is not used by the stage-1 configs")
        return ops.resnet_block(x, self)                       # one fused autograd node (ops.ResnetBlockFn)


class AttnBlock(_Nhwc):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def forward_nhwc(self, x):
        b, h, w, c = x.shape
        hn = ops.gn_swish(x, self.norm, swish=False)
        o = ops.AttnQKVFn.apply(hn, self.q.weight, self.q.bias, self.k.weight, self.k.bias, self.v.weight,
                                self.v.bias).view(b, h, w, c)
        return ops.conv2d(o, self.proj_out, residual=x)
