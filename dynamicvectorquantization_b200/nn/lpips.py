"""LPIPS perceptual loss of the stage-1 training step on the sm_100a convolution kernels.

Mirror of ``modules/losses/lpips.py`` (reference): ``LPIPS`` (:11-56), ``ScalingLayer`` (:59-66),
``NetLinLayer`` (:69-75), ``vgg16`` (:78-113), ``normalize_tensor`` / ``spatial_average`` (:116-122) with the
same constructor arguments, sub-module names and ``state_dict`` keys (``net.slice{1..5}.{torchvision
index}.{weight,bias}``, ``lin{0..4}.model.1.weight``, ``scaling_layer.{shift,scale}``).

What differs is how it is computed: the 13 VGG16 convolutions (40 GFLOP per 256x256 image, the bulk of
the loss) run as NHWC bf16 tap GEMMs on the tensor cores with the ReLU in the GEMM epilogue, max pooling
and the ReLU / pooling gradients are small CUDA kernels, and only the gradient the training step needs
(w.r.t. the reconstruction, no weight gradients - the network is frozen) is computed.  The per-level
head (channel normalisation, squared difference, dropout, 1x1 ``lin``, spatial mean) is one fused kernel per
direction (csrc/lpips.cu) that reads the two feature maps once.

Weights: like the reference, the constructor wants torchvision's pretrained VGG16 and
``modules/lpips/vgg.pth`` (the five ``lin`` heads that ship with the reference tree).  Offline, set
``B200DQ_ALLOW_RANDOM_VGG=1`` to keep a (seeded) random initialisation - meant for benchmarking and for
parity tests, which load explicit weights anyway.
"""
import os
import sys
import warnings
from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import kernels as kn
from .. import ops
from ..ops import _f32, _packed

BF16 = torch.bfloat16

# torchvision.models.vgg16().features: 'M' = MaxPool2d(2, 2); every number is a 3x3 convolution + ReLU
_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
_SLICE_ENDS = [4, 9, 16, 23, 30]                     # lpips.py:88-97


def _allow_random():
    return os.environ.get("B200DQ_ALLOW_RANDOM_VGG", "0") == "1"


# --------------------------------------------------------------------------------- autograd pieces
class _ConvReluFn(torch.autograd.Function):
    """relu(conv3x3(x) + b) on NHWC bf16, Cin % 64 == 0; ReLU fused into the GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        cout, cin = weight.shape[0], weight.shape[1]
        y = kn.conv_fwd(x, _packed(weight, "fwd"), _f32(bias), 3, 1, cout, relu=True)
        ctx.save_for_backward(x, weight, y)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        g = kn.relu_bwd(dy.contiguous(), y)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = kn.conv_dgrad(g, _packed(weight, "dgrad"), 3, 1, weight.shape[1], x.shape[1:3])
        if ctx.needs_input_grad[1]:
            dw = kn.conv_wgrad(x, g, 3, 1)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = kn.bias_grad(g)
        return dx, dw, db


class _ConvInReluFn(torch.autograd.Function):
    """First VGG layer, relu(conv3x3(image) + b) with Cin = 3: the 3x3x3 window is gathered to 64 columns and
    contracted as one GEMM tap; the gradient w.r.t. the IMAGE (what the reconstruction needs) is a 9-tap GEMM
    with the mirrored, transposed filter."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        col = kn.im2col3x3_small(x)
        dims, strs = kn.nhwc_view(col)
        y = torch.empty(nb, h, w, cout, dtype=BF16, device=x.device)
        kn.tapgemm(col, dims, strs, _packed(weight, "col_fwd"), cout, 64, [(0, 0, 0, 0, 0)], 1, y, 0,
                   (h * w * cout, w * cout, cout), w, h, nb, cout, bias=_f32(bias), relu=True)
        ctx.save_for_backward(weight, y)
        ctx.cin = cin
        return y

    @staticmethod
    def backward(ctx, dy):
        weight, y = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None
        g = kn.relu_bwd(dy.contiguous(), y)
        nb, h, w, cout = g.shape
        cin = ctx.cin
        # dx[p, ci] = sum_{r,s,co} g[p - (r-1, s-1), co] * W[co, ci, r, s]: a convolution of g with the filter
        # W'[ci, co, r, s] = W[co, ci, 2-r, 2-s], i.e. the few-output-channel form of ops.ConvOutFn.forward
        wt = weight.detach().permute(1, 0, 2, 3).flip(2, 3).contiguous()        # [cin, cout, 3, 3]
        wp = torch.zeros(16, 9 * cout, dtype=BF16, device=g.device)
        wp[:cin] = kn.pack_weight_fwd(wt)
        dims, strs = kn.nhwc_view(g)
        taps = [(0, s - 1, 0, r - 1, (r * 3 + s) * cout) for r, s in kn.TAPS_3x3]
        dx = torch.empty(nb, h, w, cin, dtype=torch.float32, device=g.device)
        kn.tapgemm(g, dims, strs, wp, 16, wp.shape[1], taps, cout // 64, dx, 0, (h * w * cin, w * cin, cin), w, h,
                   nb, cin, out_f32=True, block_n=16)
        return dx.to(BF16), None, None


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return kn.maxpool2x2(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return kn.maxpool2x2_bwd(dy.contiguous(), x)


class _HeadFn(torch.autograd.Function):
    """One LPIPS level: mean_hw sum_c w_c drop_c (f0/|f0| - f1/|f1|)^2 -> [N] fp32, fused (csrc/lpips.cu)."""

    @staticmethod
    def forward(ctx, f0, f1, w, seed, p_drop):
        f0, f1 = f0.contiguous(), f1.contiguous()
        w = w.detach().reshape(-1).float().contiguous()
        ctx.save_for_backward(f0, f1, w, seed)
        ctx.p_drop = p_drop
        return kn.lpips_head_fwd(f0, f1, w, seed, p_drop)

    @staticmethod
    def backward(ctx, g):
        f0, f1, w, seed = ctx.saved_tensors
        d0, d1 = kn.lpips_head_bwd(f0, f1, w, g.float().contiguous(), ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                   seed, ctx.p_drop)
        return d0, d1, None, None, None


# ---------------------------------------------------------------------------------------- modules
class ScalingLayer(nn.Module):
    """lpips.py:59-66."""

    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.Tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """A single linear layer which does a 1x1 conv (lpips.py:69-75)."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout(), ] if use_dropout else []
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False), ]
        self.model = nn.Sequential(*layers)


def _vgg16_features():
    """Parameter containers with torchvision's module indices (state_dict compatible)."""
    layers, cin = [], 3
    for v in _VGG16_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            cin = v
    return layers


class vgg16(torch.nn.Module):
    """lpips.py:78-113: the five feature slices of torchvision's VGG16."""

    def __init__(self, requires_grad=False, pretrained=True):
        super().__init__()
        feats = _vgg16_features()
        cached = os.path.join(torch.hub.get_dir(), "checkpoints", "vgg16-397923af.pth")
        if pretrained and _allow_random() and not os.path.exists(cached):
            warnings.warn("LPIPS: no cached VGG16 weights and B200DQ_ALLOW_RANDOM_VGG=1: random initialisation")
        elif pretrained:
            try:
                from torchvision import models
                src = models.vgg16(pretrained=True).features
                for dst, s in zip(feats, src):
                    if isinstance(dst, nn.Conv2d):
                        dst.load_state_dict(s.state_dict())
            except Exception as e:                      # no torchvision / no network / no cached weights
                if not _allow_random():
                    raise RuntimeError(
                        "LPIPS needs torchvision's pretrained VGG16 weights, which could not be loaded "
                        f"({type(e).__name__}: {e}).  Put them in the torch hub cache, or set "
                        "B200DQ_ALLOW_RANDOM_VGG=1 to run with a random initialisation (benchmarks / tests).")
                warnings.warn("LPIPS: pretrained VGG16 weights unavailable, keeping the random initialisation")
        self.N_slices = 5
        start = 0
        for k, end in enumerate(_SLICE_ENDS):
            seq = torch.nn.Sequential()
            for x in range(start, end):
                seq.add_module(str(x), feats[x])
            setattr(self, f"slice{k + 1}", seq)
            start = end
        if not requires_grad:
            for param in self.parameters():
                param.requires_grad = False

    def forward_nhwc(self, h):
        """h: NHWC bf16 image -> the five ReLU feature maps (NHWC bf16)."""
        outs = []
        for k in range(5):
            for m in getattr(self, f"slice{k + 1}"):
                if isinstance(m, nn.Conv2d):
                    if m.in_channels % 64 == 0:
                        h = _ConvReluFn.apply(h, m.weight, m.bias)
                    else:
                        h = _ConvInReluFn.apply(h, m.weight, m.bias)
                elif isinstance(m, nn.MaxPool2d):
                    h = _MaxPoolFn.apply(h)
                # nn.ReLU: fused into the convolution before it
            outs.append(h)
        return outs

    def forward(self, X):
        if not X.is_cuda:
            raise RuntimeError("vgg16 (B200) needs CUDA tensors; there is no CPU fallback")
        outs = [ops.to_nchw(o) for o in self.forward_nhwc(ops.to_nhwc(X))]
        vgg_outputs = namedtuple("VggOutputs", ["relu1_2", "relu2_2", "relu3_3", "relu4_3", "relu5_3"])
        return vgg_outputs(*outs)


def normalize_tensor(x, eps=1e-10):
    norm_factor = torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True))
    return x / (norm_factor + eps)


def spatial_average(x, keepdim=True):
    return x.mean([2, 3], keepdim=keepdim)


def _find_lin_checkpoint():
    rel = os.path.join("modules", "lpips", "vgg.pth")
    for root in [os.getcwd()] + list(sys.path):
        cand = os.path.join(root or ".", rel)
        if os.path.exists(cand):
            return cand
    return None


class LPIPS(nn.Module):
    """Learned perceptual metric (lpips.py:11-56)."""

    def __init__(self, use_dropout=True):
        super().__init__()
        self.scaling_layer = ScalingLayer()
        self.chns = [64, 128, 256, 512, 512]
        self.net = vgg16(pretrained=True, requires_grad=False)
        self.lin0 = NetLinLayer(self.chns[0], use_dropout=use_dropout)
        self.lin1 = NetLinLayer(self.chns[1], use_dropout=use_dropout)
        self.lin2 = NetLinLayer(self.chns[2], use_dropout=use_dropout)
        self.lin3 = NetLinLayer(self.chns[3], use_dropout=use_dropout)
        self.lin4 = NetLinLayer(self.chns[4], use_dropout=use_dropout)
        self.load_from_pretrained()
        for param in self.parameters():
            param.requires_grad = False

    def load_from_pretrained(self, name="vgg_lpips"):
        ckpt = _find_lin_checkpoint()
        if ckpt is None:
            if not _allow_random():
                raise RuntimeError("LPIPS: modules/lpips/vgg.pth (the lin heads shipped with the reference tree) not "
                                   "found under the working directory or sys.path; set B200DQ_ALLOW_RANDOM_VGG=1 to "
                                   "run with seeded non-negative heads (benchmarks / tests).")
            g = torch.Generator().manual_seed(0)
            with torch.no_grad():
                for k, c in enumerate(self.chns):
                    getattr(self, f"lin{k}").model[-1].weight.copy_(torch.rand(1, c, 1, 1, generator=g) * (2.0 / c))
            return
        self.load_state_dict(torch.load(ckpt, map_location=torch.device("cpu")), strict=False)
        print("loaded pretrained LPIPS loss from {}".format(ckpt))

    @classmethod
    def from_pretrained(cls, name="vgg_lpips"):
        if name != "vgg_lpips":
            raise NotImplementedError
        return cls()

    def _head(self, k, f0, f1):
        """Level k: mean_hw sum_c w_c (f0/|f0| - f1/|f1|)^2 on NHWC features -> [N] (fp32).  The Dropout in front
        of the lin head is live whenever the module is in training mode, as in the reference (LPIPS().eval() at
        vqperceptual_multidisc.py:74 does not survive the LightningModule's .train())."""
        lin = getattr(self, f"lin{k}").model
        seed, p_drop = None, 0.0
        for m in lin:
            if isinstance(m, nn.Dropout) and m.training and m.p > 0:
                # drawn on the device so that a captured CUDA graph gets a fresh mask at every replay
                seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=f0.device)
                p_drop = float(m.p)
        return _HeadFn.apply(f0, f1, lin[-1].weight, seed, p_drop)

    def forward(self, input, target):
        if not input.is_cuda:
            raise RuntimeError("LPIPS (B200) needs CUDA tensors; there is no CPU fallback")
        in0, in1 = self.scaling_layer(input), self.scaling_layer(target)
        # the side that needs no gradient (the real image in the training step) runs without a graph
        if in0.requires_grad:
            outs0 = self.net.forward_nhwc(ops.to_nhwc(in0))
        else:
            with torch.no_grad():
                outs0 = self.net.forward_nhwc(ops.to_nhwc(in0))
        if in1.requires_grad:
            outs1 = self.net.forward_nhwc(ops.to_nhwc(in1))
        else:
            with torch.no_grad():
                outs1 = self.net.forward_nhwc(ops.to_nhwc(in1))
        val = self._head(0, outs0[0], outs1[0])
        for k in range(1, len(self.chns)):
            val = val + self._head(k, outs0[k], outs1[k])
        return val.reshape(-1, 1, 1, 1)
