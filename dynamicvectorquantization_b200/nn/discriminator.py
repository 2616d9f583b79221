"""PatchGAN discriminator of the stage-1 training step on the sm_100a kernels.

Mirror of ``modules/discriminator/model.py`` (reference): ``weights_init`` (:8-14) and
``NLayerDiscriminator`` (:17-67) with the same constructor arguments and ``state_dict`` keys
(``main.{0,2,5,8,11}.weight``, BatchNorm at ``main.{3,6,9}``); ``main`` is the same ``nn.Sequential`` of
``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.LeakyReLU`` modules, used as parameter containers.

What differs is how ``forward`` is computed (NHWC bf16, fp32 accumulation, no cuDNN / ATen convolution or
batch-norm kernel):
* the 4x4 stride-2 convolutions are 16-tap tensor-core GEMMs over the stride-2 parity view of the input
  (``kernels.conv4x4_*``; data gradient = four 4-tap GEMMs, one per input parity class; weight gradient = the
  split-K pixel-contraction GEMM), the 4x4 stride-1 stage the same with plain shifted boxes;
* the 3-channel stem gathers its 4x4x3 window to 64 columns (``im2col_window``) and runs one GEMM tap with the
  LeakyReLU in the epilogue; its gradient w.r.t. the IMAGE (the path from the generator loss to the decoder)
  is four 4-tap GEMMs with 3 valid output channels;
* the one-channel head is a 16-tap GEMM with one valid output column; its backward gathers the 4x4 window of
  dy once and reuses it for dX (one tap) and dW;
* training-mode ``BatchNorm2d`` + ``LeakyReLU(0.2)`` are the GroupNorm kernels on the tensor seen as one image
  with one group per channel (statistics in fp32 / fp64, running statistics updated like ``nn.BatchNorm2d``:
  momentum 0.1, unbiased variance).
``use_actnorm=True`` (no stage-1 config uses it) keeps the reference's ActNorm module on PyTorch ops.
"""
import torch
import torch.nn as nn

from .. import kernels as kn
from .. import ops
from ..ops import _f32, _packed

BF16 = torch.bfloat16
LRELU = 2                      # activation code of the norm / GEMM-epilogue kernels: LeakyReLU(0.2)


def weights_init(m):
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm") != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


# --------------------------------------------------------------------------------- autograd pieces
class _StemFn(torch.autograd.Function):
    """LeakyReLU(conv4x4 stride 2 pad 1 (image) + b), Cin*16 <= 64 (discriminator/model.py:37)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        ho, wo = h // 2, w // 2
        col = kn.im2col_window(x, 4, 2, 1, -1, (ho, wo))
        dims, strs = kn.nhwc_view(col)
        y = torch.empty(nb, ho, wo, cout, dtype=BF16, device=x.device)
        kn.tapgemm(col, dims, strs, _packed(weight, "col_fwd"), cout, 64, [(0, 0, 0, 0, 0)], 1, y, 0,
                   (ho * wo * cout, wo * cout, cout), wo, ho, nb, cout, bias=_f32(bias), relu=LRELU)
        ctx.save_for_backward(x, weight, y)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        g = kn.lrelu_bwd(dy.contiguous(), y)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dx[2i+ph, 2j+pw, ci] = sum over the taps r = ph+1 (mod 2), s = pw+1 (mod 2) of g[i+dh, j+dw, :] . W[:, ci, r, s]
            wp = torch.zeros(16, 16 * cout, dtype=BF16, device=g.device)
            wp[:cin] = _packed(weight, "dgrad")
            dx = torch.empty(nb, h, w, cin, dtype=BF16, device=g.device)
            dims, strs = kn.nhwc_view(g)
            rsel = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}
            for ph in (0, 1):
                for pw in (0, 1):
                    taps = [(0, dwo, 0, dho, (r * 4 + s) * cout) for r, dho in rsel[ph] for s, dwo in rsel[pw]]
                    kn.tapgemm(g, dims, strs, wp, 16, wp.shape[1], taps, cout // 64, dx, (ph * w + pw) * cin,
                               (h * w * cin, 2 * w * cin, 2 * cin), w // 2, h // 2, nb, cin, block_n=16)
        if ctx.needs_input_grad[1]:
            col = kn.im2col_window(x, 4, 2, 1, -1, (h // 2, w // 2))
            dwc = ops._col_wgrad(g, col, cout)                                    # [cout, 64], column = t*cin + c
            dw = dwc[:, :16 * cin].reshape(cout, 4, 4, cin).permute(0, 3, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = kn.bias_grad(g)
        return dx, dw, db


class _Conv4x4Fn(torch.autograd.Function):
    """conv4x4 (stride 2 | 1, pad 1) of an NHWC bf16 tensor with Cin, Cout multiples of 64 (+ bias)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        y = kn.conv4x4_fwd(x, _packed(weight, "fwd"), _f32(bias), stride, weight.shape[0])
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.has_bias = stride, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = kn.conv4x4_dgrad(dy, _packed(weight, "dgrad"), ctx.stride, weight.shape[1], x.shape[1:3])
        if ctx.needs_input_grad[1]:
            dw = kn.conv4x4_wgrad(x, dy, ctx.stride)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = kn.bias_grad(dy)
        return dx, dw, db, None


class _HeadFn(torch.autograd.Function):
    """conv4x4 stride 1 pad 1 to ONE channel (discriminator/model.py:66), fp32 NHWC output [N,H-1,W-1,1]."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = kn.conv4x4_fwd(x, _packed(weight, "fwd_pad16"), _f32(bias), 1, weight.shape[0], out_f32=True, block_n=16)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        nb, h, w, cin = x.shape
        cout = weight.shape[0]
        # dy of a mean-type loss is the SAME value at every position (e.g. -1/numel): rounding it to bf16 once would
        # bias every discriminator gradient by up to 2^-9.  It is carried as two bf16 terms (hi + lo, 16 mantissa
        # bits) in two window channels that meet the same filter weights.
        d32 = dy.float().contiguous()
        hi = d32.to(BF16)
        d2 = torch.cat([hi, (d32 - hi.float()).to(BF16)], dim=-1).contiguous()    # [N,H',W',2*cout]
        # window of dy each input pixel sees: col[n,ih,iw,(r*4+s)*2*cout + (c*cout+co)] = d_c[n, ih+1-r, iw+1-s, co]
        colf = kn.im2col_window(d2, 4, 1, -1, 1, (h, w))
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(nb, h, w, cin, dtype=BF16, device=x.device)
            dims, strs = kn.nhwc_view(colf)
            kn.tapgemm(colf, dims, strs, _packed(weight, "col_dgrad_x2"), cin, 64, [(0, 0, 0, 0, 0)], 1, dx, 0,
                       (h * w * cin, w * cin, cin), w, h, nb, cin)
        if ctx.needs_input_grad[1]:
            dwc = ops._col_wgrad(x, colf, cin)                                    # [cin, 64]: [ci, t*2*cout + c*cout + co]
            dwc = dwc[:, :32 * cout].reshape(cin, 16, 2, cout).sum(2)             # hi + lo
            dw = dwc.reshape(cin, 4, 4, cout).permute(3, 0, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum((0, 1, 2))
        return dx, dw, db


class _BatchNormActFn(torch.autograd.Function):
    """LeakyReLU(BatchNorm2d(x)) on NHWC bf16.  stats [1,C,2] = (mean, rstd) - of the batch in training mode, from the
    running statistics in evaluation mode (then they are constants of the backward)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, stats, batch_stats):
        g, b = _f32(gamma), _f32(beta)
        ctx.save_for_backward(x, stats, g, b)
        ctx.batch_stats = batch_stats
        return kn.bn_apply(x, stats, g, b, LRELU)

    @staticmethod
    def backward(ctx, dy):
        x, stats, g, b = ctx.saved_tensors
        dx, dg, db = kn.bn_bwd(dy.contiguous(), x, stats, g, b, LRELU, batch_stats=ctx.batch_stats)
        return dx, dg, db, None, None


def _batchnorm_lrelu(x, bn):
    """x NHWC bf16; bn: the nn.BatchNorm2d parameter / buffer container (updated like nn.BatchNorm2d.forward)."""
    use_batch = bn.training or not bn.track_running_stats
    if use_batch:
        stats = kn.bn_stats(x.detach(), eps=bn.eps)
        if bn.training and bn.track_running_stats:
            with torch.no_grad():
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                cnt = x.numel() // x.shape[-1]
                mean, rstd = stats[0, :, 0], stats[0, :, 1]
                var = (1.0 / (rstd * rstd) - bn.eps).clamp_min_(0.0)
                bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
                bn.running_var.mul_(1 - m).add_(var, alpha=m * cnt / max(cnt - 1, 1))
    else:
        stats = torch.stack([bn.running_mean.float(), (bn.running_var.float() + bn.eps).rsqrt()], dim=1).unsqueeze(0)
        stats = stats.contiguous()
    return _BatchNormActFn.apply(x, bn.weight, bn.bias, stats, use_batch)


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
        self._on_kernels = (not use_actnorm and input_nc * 16 <= 64 and ndf % 64 == 0)

    def forward(self, input):
        if not self._on_kernels:
            return self.main(input)                       # ActNorm variant / odd widths: the reference's own module graph
        if not input.is_cuda:
            raise RuntimeError("NLayerDiscriminator (B200) needs CUDA tensors; there is no CPU fallback")
        mods = list(self.main)
        h = _StemFn.apply(ops.to_nhwc(input), mods[0].weight, mods[0].bias)
        i = 2
        while i + 2 < len(mods):                          # conv -> BatchNorm -> LeakyReLU stages
            conv, bn = mods[i], mods[i + 1]
            h = _Conv4x4Fn.apply(h, conv.weight, conv.bias, conv.stride[0])
            h = _batchnorm_lrelu(h, bn)
            i += 3
        head = mods[i]
        y = _HeadFn.apply(h, head.weight, head.bias)      # [N, H', W', 1] fp32
        return y.permute(0, 3, 1, 2)                      # NCHW view of a one-channel map (same memory order)
