"""Positional decoder of DQ-VAE on the sm_100a kernels.

Mirrors ``modules/dynamic_modules/DecoderPositional.py:13-145`` and ``fourier_embedding.py:5-55``
(constructor arguments, sub-module names and order, ``forward(h, grain_indices)``; it must expose
``conv_out.weight`` for the reference loss' adaptive weight).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .blocks import AttnBlock, Normalize, ResnetBlock, Upsample, _require_cuda


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def convert_to_coord_format(b, h, w, device="cpu", integer_values=False):
    if integer_values:
        xs = torch.arange(w, dtype=torch.float, device=device)
        ys = torch.arange(h, dtype=torch.float, device=device)
    else:
        xs = torch.linspace(-1, 1, w, device=device)
        ys = torch.linspace(-1, 1, h, device=device)
    x_channel = xs.view(1, 1, 1, -1).repeat(b, 1, w, 1)
    y_channel = ys.view(1, 1, -1, 1).repeat(b, 1, 1, h)
    return torch.cat((x_channel, y_channel), dim=1)


class ConLinear(nn.Module):
    def __init__(self, ch_in, ch_out, is_first=False, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(ch_in, ch_out, kernel_size=1, padding=0, bias=bias)
        bound = np.sqrt(9 / ch_in) if is_first else np.sqrt(3 / ch_in)
        nn.init.uniform_(self.conv.weight, -bound, bound)

    def forward(self, x):
        # 1x1 conv over a 2-channel coordinate grid == two broadcast multiply-adds (no cuDNN)
        w = self.conv.weight[:, :, 0, 0]                       # [out, in]
        y = torch.einsum("bihw,oi->bohw", x, w) if x.shape[1] > 4 else sum(
            x[:, i:i + 1] * w[:, i].view(1, -1, 1, 1) for i in range(x.shape[1]))
        if self.conv.bias is not None:
            y = y + self.conv.bias.view(1, -1, 1, 1)
        return y


class SinActivation(nn.Module):
    def forward(self, x):
        return torch.sin(x)


class LFF(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.ffm = ConLinear(2, hidden_size, is_first=True)
        self.activation = SinActivation()

    def forward(self, x):
        return self.activation(self.ffm(x))


class FourierPositionEmbedding(nn.Module):
    def __init__(self, coord_size, hidden_size, integer_values=False):
        super().__init__()
        self.coord = convert_to_coord_format(1, coord_size, coord_size, "cpu", integer_values)
        self.lff = LFF(hidden_size)

    def bias(self, device):
        dev = torch.device(device)
        cached = getattr(self, "_coord_dev", None)
        if cached is None or cached.device != dev:             # one H2D copy per device, not per step
            cached = self.coord.to(dev)
            self._coord_dev = cached
        return self.lff(cached)                                # [1, C, h, w]

    def forward(self, x):
        return x + self.bias(x.device)


class PositionEmbedding2DLearned(nn.Module):
    def __init__(self, n_row, feats_dim, n_col=None):
        super().__init__()
        n_col = n_col if n_col is not None else n_row
        self.row_embed = nn.Embedding(n_row, feats_dim)
        self.col_embed = nn.Embedding(n_col, feats_dim)
        self.reset_parameters()

    def reset_parameters(self):
        trunc_normal_(self.row_embed.weight)
        trunc_normal_(self.col_embed.weight)

    def bias(self, h, w):
        x_emb = self.col_embed.weight[:w].unsqueeze(0)         # [1, w, C]
        y_emb = self.row_embed.weight[:h].unsqueeze(1)         # [h, 1, C]
        return (x_emb + y_emb).permute(2, 0, 1).unsqueeze(0)   # [1, C, h, w]

    def forward(self, x):
        h, w = x.shape[-2:]
        pos = self.bias(h, w)
        if x.dim() == 5:
            pos = pos.unsqueeze(-3)
        return x + pos


class Decoder(nn.Module):
    def __init__(self, ch, in_ch, out_ch, ch_mult, num_res_blocks, resolution, attn_resolutions,
                 dropout=0.0, resamp_with_conv=True, give_pre_end=False, latent_size=32, window_size=2,
                 position_type="relative"):
        super().__init__()
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_ch = in_ch
        self.temb_ch = 0
        self.ch = ch
        self.give_pre_end = give_pre_end
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, in_ch, curr_res, curr_res)
        self.conv_in = torch.nn.Conv2d(in_ch, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block = block
            up.attn = attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = torch.nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self.position_type = position_type
        if self.position_type == "learned":
            self.position_bias = PositionEmbedding2DLearned(n_row=latent_size, feats_dim=in_ch)
        elif self.position_type == "fourier":
            self.position_bias = FourierPositionEmbedding(coord_size=latent_size, hidden_size=in_ch)
        elif self.position_type == "fourier+learned":
            self.position_bias_fourier = FourierPositionEmbedding(coord_size=latent_size, hidden_size=in_ch)
            self.position_bias_learned = PositionEmbedding2DLearned(n_row=latent_size, feats_dim=in_ch)
        else:
            raise NotImplementedError(f"position_type {position_type!r} is not used by the stage-1 configs")

    def position_bias_nchw(self, h, w, device):
        """Sum of the configured position embeddings, [1, C, h, w] fp32 (autograd-tracked)."""
        if self.position_type == "fourier":
            return self.position_bias.bias(device)
        if self.position_type == "learned":
            return self.position_bias.bias(h, w)
        return self.position_bias_fourier.bias(device) + self.position_bias_learned.bias(h, w)

    def forward_nhwc(self, h):
        """h NHWC bf16 with the position bias already added."""
        h = ops.conv2d(h, self.conv_in)
        h = self.mid.block_1.forward_nhwc(h)
        h = self.mid.attn_1.forward_nhwc(h)
        h = self.mid.block_2.forward_nhwc(h)
        for i_level in reversed(range(self.num_resolutions)):
            lvl = self.up[i_level]
            for i_block in range(self.num_res_blocks + 1):
                h = lvl.block[i_block].forward_nhwc(h)
                if len(lvl.attn) > 0:
                    h = lvl.attn[i_block].forward_nhwc(h)
            if i_level != 0:
                h = lvl.upsample.forward_nhwc(h)
        if self.give_pre_end:
            return h
        h = ops.gn_swish(h, self.norm_out)
        return ops.ConvOutFn.apply(h, self.conv_out.weight, self.conv_out.bias)   # NHWC fp32

    def forward(self, h, grain_indices=None):
        _require_cuda(h)
        h = h + self.position_bias_nchw(h.shape[-2], h.shape[-1], h.device)
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(h)))
