"""Entropy thresholds for the fixed-entropy router (second half).

Mirror of ``scripts/tools/calculate_entropy_thresholds.py`` (reference): the per-patch grey-level entropy of
every image of a data set (its ``Entropy`` module, :27-79 - the same soft histogram as the model's but with the
32 bins on [0, 1]), sorted, and the 99 percentile thresholds ``sorted[(size * i) // 100]`` written as the JSON
that ``DualGrainFixedEntropyRouter`` reads (:95-117).  Here the entropies come from the fused CUDA kernel
(``csrc/entropy.cu``: one read of the image, no [B*patches, pixels, bins] intermediate) and stay on the device
until the final sort.

    acc = EntropyThresholds(patch_size=16, image_size=256)
    for batch in loader:                      # [B,3,H,W] fp32 CUDA, same value range the tool was given
        acc.update(batch)
    acc.save("scripts/tools/thresholds/entropy_thresholds_imagenet_train_patch-16.json")
"""
import json

import torch

from .. import kernels as kn


class EntropyThresholds:
    def __init__(self, patch_size=16, image_size=256, bins_lo=0.0, bins_hi=1.0, num_bins=32, sigma=0.01):
        self.psize = patch_size
        self.image_size = image_size
        self.lo, self.hi, self.num_bins, self.sigma = bins_lo, bins_hi, num_bins, sigma
        self._bins = None
        self._chunks = []

    @torch.no_grad()
    def entropy(self, images):
        """[B,3,H,W] fp32 CUDA -> [B, H/p, W/p] patch entropies (calculate_entropy_thresholds.py:65-79)."""
        if not images.is_cuda:
            raise RuntimeError("EntropyThresholds (B200) needs CUDA tensors; there is no CPU fallback")
        assert images.shape[-1] == self.image_size and images.shape[-2] == self.image_size
        if self._bins is None or self._bins.device != images.device:
            self._bins = torch.linspace(self.lo, self.hi, self.num_bins).to(images.device)
        return kn.patch_entropy(images.float(), self._bins, self.psize, self.sigma)

    @torch.no_grad()
    def update(self, images):
        self._chunks.append(self.entropy(images).reshape(-1))

    @torch.no_grad()
    def thresholds(self):
        """{"1": t1, ..., "99": t99} with t_i = sorted[(size * i) // 100] (:108-116)."""
        ent = torch.sort(torch.cat(self._chunks)).values
        size = ent.numel()
        pos = torch.tensor([(size * (i + 1)) // 100 for i in range(99)], device=ent.device)
        vals = ent[pos].cpu().tolist()
        return {str(i + 1): float(v) for i, v in enumerate(vals)}

    def save(self, path):
        with open(path, "w") as f:
            json.dump(self.thresholds(), f)
