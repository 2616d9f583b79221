"""Grain routers and budget losses (tiny, fp32, stay in PyTorch - SURVEY.md 2a #8).

Mirrors ``modules/dynamic_modules/RouterDual.py`` (:6-57), ``RouterTriple.py`` (:6-55) and
``budget.py`` (:4-59) of the reference: same class names, constructor arguments, parameter names.
"""
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


class _AvgPoolFn(torch.autograd.Function):
    """F.avg_pool2d(x, k, k) with its gradient written as a scaled broadcast (ATen's avg_pool2d_backward kernel takes
    125 us for the [32,256,32,32] router feature; a multiply and a strided copy take 20).  Same values: the gradient
    of a k x k mean is g / k^2 in every window position, and 1 / k^2 is a power of two for k = 2, 4."""

    @staticmethod
    def forward(ctx, x, k):
        ctx.k = k
        return F.avg_pool2d(x, k, k)

    @staticmethod
    def backward(ctx, g):
        k = ctx.k
        b, c, h, w = g.shape
        g = g * (1.0 / (k * k))
        return g[:, :, :, None, :, None].expand(b, c, h, k, w, k).reshape(b, c, h * k, w * k), None


class _AvgPool(nn.Module):
    """nn.AvgPool2d(k, k) (no parameters, no state_dict entries) on _AvgPoolFn."""

    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, x):
        if x.shape[-1] % self.k or x.shape[-2] % self.k:
            return F.avg_pool2d(x, self.k, self.k)
        return _AvgPoolFn.apply(x, self.k)


def _gate_mlp(width, n_out, gate_type):
    if gate_type == "1layer-fc":
        return nn.Linear(width, n_out)
    if gate_type == "2layer-fc-SiLu":
        return nn.Sequential(nn.Linear(width, width), nn.SiLU(inplace=True), nn.Linear(width, n_out))
    if gate_type == "2layer-fc-ReLu":
        return nn.Sequential(nn.Linear(width, width), nn.ReLU(inplace=True), nn.Linear(width, n_out))
    raise NotImplementedError(gate_type)


def _feature_norm(normalization_type, num_channels):
    if normalization_type == "none":
        return nn.Identity()
    if "group" in normalization_type:
        groups = int(normalization_type.split("-")[-1])
        return nn.GroupNorm(num_groups=groups, num_channels=num_channels, eps=1e-6, affine=True)
    raise NotImplementedError(normalization_type)


class DualGrainFeatureRouter(nn.Module):
    def __init__(self, num_channels, normalization_type="none", gate_type="1layer-fc"):
        super().__init__()
        self.gate_pool = _AvgPool(2)
        self.gate_type = gate_type
        if gate_type not in ("1layer-fc", "2layer-fc-SiLu"):
            raise NotImplementedError()
        self.gate = _gate_mlp(num_channels * 2, 2, gate_type)
        self.num_splits = 2
        self.normalization_type = normalization_type
        self.feature_norm_fine = _feature_norm(normalization_type, num_channels)
        self.feature_norm_coarse = _feature_norm(normalization_type, num_channels)

    def forward(self, h_fine, h_coarse, entropy=None):
        h_fine = self.feature_norm_fine(h_fine)
        h_coarse = self.feature_norm_coarse(h_coarse)
        feats = torch.cat([h_coarse, self.gate_pool(h_fine)], dim=1).permute(0, 2, 3, 1)
        return self.gate(feats)                                   # [B, h, w, 2]


class DualGrainFixedEntropyRouter(nn.Module):
    def __init__(self, json_path, fine_grain_ratito):
        super().__init__()
        with open(json_path, "r", encoding="utf-8") as f:
            content = json.load(f)
        self.fine_grain_threshold = content["{}".format(str(int(100 - fine_grain_ratito * 100)))]

    def forward(self, h_fine=None, h_coarse=None, entropy=None):
        fine = (entropy > self.fine_grain_threshold).bool().long().unsqueeze(-1)
        coarse = (entropy <= self.fine_grain_threshold).bool().long().unsqueeze(-1)
        return torch.cat([coarse, fine], dim=-1)


class TripleGrainFeatureRouter(nn.Module):
    def __init__(self, num_channels, normalization_type="none", gate_type="1layer-fc"):
        super().__init__()
        self.gate_median_pool = _AvgPool(2)
        self.gate_fine_pool = _AvgPool(4)
        self.num_splits = 3
        self.gate_type = gate_type
        self.gate = _gate_mlp(num_channels * 3, 3, gate_type)
        self.normalization_type = normalization_type
        self.feature_norm_fine = _feature_norm(normalization_type, num_channels)
        self.feature_norm_median = _feature_norm(normalization_type, num_channels)
        self.feature_norm_coarse = _feature_norm(normalization_type, num_channels)

    def forward(self, h_fine, h_median, h_coarse, entropy=None):
        h_fine = self.feature_norm_fine(h_fine)
        h_median = self.feature_norm_median(h_median)
        h_coarse = self.feature_norm_coarse(h_coarse)
        feats = torch.cat([h_coarse, self.gate_median_pool(h_median), self.gate_fine_pool(h_fine)], dim=1)
        return self.gate(feats.permute(0, 2, 3, 1))


class BudgetConstraint_RatioMSE_DualGrain(nn.Module):
    def __init__(self, target_ratio=0., gamma=1.0, min_grain_size=8, max_grain_size=16, calculate_all=True):
        super().__init__()
        self.target_ratio = target_ratio
        self.gamma = gamma
        self.calculate_all = calculate_all
        self.loss = nn.MSELoss()
        self.const = min_grain_size * min_grain_size
        self.max_const = max_grain_size * max_grain_size - self.const

    def forward(self, gate):
        # gate [B, 2, h, w]: channel 0 = coarse (1 code), channel 1 = fine (4 codes)
        used = (1.0 * gate[:, 0] + 4.0 * gate[:, 1]).sum() / gate.size(0) - self.const
        ratio = used / self.max_const
        target = self.target_ratio * torch.ones_like(ratio)
        if self.calculate_all:
            last = self.gamma * self.loss(1 - ratio, 1 - target)
            return last + last                                    # reference quirk (budget.py:24-26)
        return self.gamma * self.loss(ratio, target)


class BudgetConstraint_NormedSeperateRatioMSE_TripleGrain(nn.Module):
    def __init__(self, target_fine_ratio=0., target_median_ratio=0., gamma=1.0, min_grain_size=8,
                 median_grain_size=16, max_grain_size=32):
        super().__init__()
        assert target_fine_ratio + target_median_ratio <= 1.0
        self.target_fine_ratio = target_fine_ratio
        self.target_median_ratio = target_median_ratio
        self.gamma = gamma
        self.loss = nn.MSELoss()
        self.min_const = min_grain_size * min_grain_size
        self.median_const = median_grain_size * median_grain_size - self.min_const
        self.max_const = max_grain_size * max_grain_size - self.min_const

    def forward(self, gate):
        # gate [B, 3, h, w]: coarse, median (x4), fine (x16); the extra 1.0 terms are the
        # reference's compensation terms (budget.py:46,53)
        n = gate.size(0)
        med = (gate[:, 0] + 4.0 * gate[:, 1] + gate[:, 2]).sum() / n - self.min_const
        r_med = med / self.median_const
        loss_med = self.loss(r_med, self.target_median_ratio * torch.ones_like(r_med))
        fine = (gate[:, 0] + 16.0 * gate[:, 2] + gate[:, 1]).sum() / n - self.min_const
        r_fine = fine / self.max_const
        loss_fine = self.gamma * self.loss(r_fine, self.target_fine_ratio * torch.ones_like(r_fine))
        return loss_fine + loss_med
