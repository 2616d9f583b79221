"""Dual- and triple-grain encoders of DQ-VAE on the sm_100a kernels.

Mirrors ``modules/dynamic_modules/EncoderDual.py:15-156`` and ``EncoderTriple.py`` of the
reference: constructor arguments, sub-module names / creation order (=> identical ``state_dict``
keys and seeded default init) and the returned dict.  The convolutional trunk and heads run
NHWC bf16 through the C-ABI kernels; routing, grain merging and the code mask are a few tiny
fp32 tensor ops (kept in PyTorch together with their RNG: ``F.gumbel_softmax``).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .blocks import AttnBlock, Downsample, Normalize, ResnetBlock, _require_cuda

try:  # the reference derives from LightningModule (EncoderDual.py:10,15)
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pytorch_lightning is not installed in this image
    _Base = nn.Module

try:
    from utils.utils import instantiate_from_config  # reference's plugin loader when on sys.path
except Exception:
    from ..config import instantiate_from_config


class _Head(nn.Module):
    pass


def _make_mid(block_in, temb_ch, dropout):
    mid = nn.Module()
    mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=temb_ch, dropout=dropout)
    mid.attn_1 = AttnBlock(block_in)
    mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=temb_ch, dropout=dropout)
    return mid


class _GrainEncoderBase(_Base):
    def _build_trunk(self, ch, ch_mult, num_res_blocks, attn_resolutions, dropout, resamp_with_conv,
                     in_channels, resolution):
        self.ch = ch
        self.temb_ch = 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution = resolution
        self.in_channels = in_channels
        self.conv_in = torch.nn.Conv2d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block = nn.ModuleList()
            attn = nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out,
                                         temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block = block
            down.attn = attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        return block_in

    def _trunk(self, x):
        """x NCHW fp32 -> list of NHWC bf16 features at the end of each level (before down-sampling)."""
        _require_cuda(x)
        assert x.shape[2] == x.shape[3] == self.resolution, "{}, {}, {}".format(x.shape[2], x.shape[3], self.resolution)
        h = ops.ConvInFn.apply(ops.to_nhwc(x), self.conv_in.weight, self.conv_in.bias)
        feats = []
        for i_level in range(self.num_resolutions):
            lvl = self.down[i_level]
            for i_block in range(self.num_res_blocks):
                h = lvl.block[i_block].forward_nhwc(h)
                if len(lvl.attn) > 0:
                    h = lvl.attn[i_block].forward_nhwc(h)
            feats.append(h)
            if i_level != self.num_resolutions - 1:
                h = lvl.downsample.forward_nhwc(h)
        return feats

    @staticmethod
    def _head(h, mid, norm_out, conv_out):
        h = mid.block_1.forward_nhwc(h)
        h = mid.attn_1.forward_nhwc(h)
        h = mid.block_2.forward_nhwc(h)
        h = ops.gn_swish(h, norm_out)
        return ops.to_nchw(ops.conv2d(h, conv_out))           # small: [B, z, 32|16|8, .] fp32 NCHW


class DualGrainEncoder(_GrainEncoderBase):
    def __init__(self, *, ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, router_config=None,
                 update_router=True, **ignore_kwargs):
        super().__init__()
        block_in = self._build_trunk(ch, ch_mult, num_res_blocks, attn_resolutions, dropout,
                                     resamp_with_conv, in_channels, resolution)
        self.mid_coarse = _make_mid(block_in, self.temb_ch, dropout)
        self.norm_out_coarse = Normalize(block_in)
        self.conv_out_coarse = torch.nn.Conv2d(block_in, z_channels, kernel_size=3, stride=1, padding=1)
        block_in_fine = block_in // (ch_mult[-1] // ch_mult[-2])
        self.mid_fine = _make_mid(block_in_fine, self.temb_ch, dropout)
        self.norm_out_fine = Normalize(block_in_fine)
        self.conv_out_fine = torch.nn.Conv2d(block_in_fine, z_channels, kernel_size=3, stride=1, padding=1)
        self.router = instantiate_from_config(router_config)
        self.update_router = update_router

    def forward(self, x, x_entropy):
        feats = self._trunk(x)
        h_coarse = self._head(feats[-1], self.mid_coarse, self.norm_out_coarse, self.conv_out_coarse)
        h_fine = self._head(feats[-2], self.mid_fine, self.norm_out_fine, self.conv_out_fine)

        # dynamic routing (EncoderDual.py:130-149): 0 = coarse, 1 = fine
        gate = self.router(h_fine=h_fine, h_coarse=h_coarse, entropy=x_entropy)
        if self.update_router and self.training:
            gate = F.gumbel_softmax(gate, dim=-1, hard=True)
        gate = gate.permute(0, 3, 1, 2)
        indices = gate.argmax(dim=1)

        up = h_coarse.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
        idx_rep = indices.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2).unsqueeze(1)
        h_dual = torch.where(idx_rep == 0, up, h_fine)
        if self.update_router and self.training:
            gate_grad = gate.max(dim=1, keepdim=True)[0]
            h_dual = h_dual * gate_grad.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
        codebook_mask = torch.where(idx_rep == 0, 0.25, 1.0).to(h_dual.dtype)
        return {"h_dual": h_dual, "indices": indices, "codebook_mask": codebook_mask, "gate": gate}


class TripleGrainEncoder(_GrainEncoderBase):
    """EncoderTriple.py: three heads (fine 32x32, median 16x16, coarse 8x8)."""

    def __init__(self, *, ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, router_config=None,
                 **ignore_kwargs):
        super().__init__()
        block_in = self._build_trunk(ch, ch_mult, num_res_blocks, attn_resolutions, dropout,
                                     resamp_with_conv, in_channels, resolution)
        self.mid_coarse = _make_mid(block_in, self.temb_ch, dropout)
        self.norm_out_coarse = Normalize(block_in)
        self.conv_out_coarse = torch.nn.Conv2d(block_in, z_channels, kernel_size=3, stride=1, padding=1)
        block_in_median = block_in // (ch_mult[-1] // ch_mult[-2])
        self.mid_median = _make_mid(block_in_median, self.temb_ch, dropout)
        self.norm_out_median = Normalize(block_in_median)
        self.conv_out_median = torch.nn.Conv2d(block_in_median, z_channels, kernel_size=3, stride=1, padding=1)
        block_in_fine = block_in_median // (ch_mult[-2] // ch_mult[-3])
        self.mid_fine = _make_mid(block_in_fine, self.temb_ch, dropout)
        self.norm_out_fine = Normalize(block_in_fine)
        self.conv_out_fine = torch.nn.Conv2d(block_in_fine, z_channels, kernel_size=3, stride=1, padding=1)
        self.router = instantiate_from_config(router_config)

    def forward(self, x, x_entropy=None):
        feats = self._trunk(x)
        h_coarse = self._head(feats[-1], self.mid_coarse, self.norm_out_coarse, self.conv_out_coarse)
        h_median = self._head(feats[-2], self.mid_median, self.norm_out_median, self.conv_out_median)
        h_fine = self._head(feats[-3], self.mid_fine, self.norm_out_fine, self.conv_out_fine)

        gate = self.router(h_fine=h_fine, h_median=h_median, h_coarse=h_coarse, entropy=x_entropy)
        if self.training:
            gate = F.gumbel_softmax(gate, tau=1, dim=-1, hard=True)
        gate = gate.permute(0, 3, 1, 2)
        indices = gate.argmax(dim=1)

        up_c = h_coarse.repeat_interleave(4, dim=-1).repeat_interleave(4, dim=-2)
        up_m = h_median.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2)
        idx_rep = indices.repeat_interleave(4, dim=-1).repeat_interleave(4, dim=-2).unsqueeze(1)
        # 0 coarse, 1 median, 2 fine
        h_triple = torch.where(idx_rep == 0, up_c, torch.where(idx_rep == 1, up_m, h_fine))
        if self.training:
            gate_grad = gate.max(dim=1, keepdim=True)[0]
            h_triple = h_triple * gate_grad.repeat_interleave(4, dim=-1).repeat_interleave(4, dim=-2)
        one = torch.ones((), dtype=h_triple.dtype, device=h_triple.device)
        codebook_mask = torch.where(idx_rep == 0, 0.0625 * one, torch.where(idx_rep == 1, 0.25 * one, one))
        return {"h_triple": h_triple, "indices": indices, "codebook_mask": codebook_mask, "gate": gate}
