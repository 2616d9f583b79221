"""Dual-grain code permuter of the stage-2 tokenisation path.

Mirror of ``modules/dynamic_modules/permuter.py:7-132`` (reference): same constructor arguments,
``forward(indices, grain_indices) -> dict`` of padded coarse / fine content, position and segment
sequences, and ``forward_back(...) -> [B, fine_hw, fine_hw]`` code map.  Both directions run as one
CUDA kernel per call (csrc/permuter.cu) instead of per-sample Python masking / per-element loops; the
only host synchronisation is reading the two longest sequence lengths that size the padded outputs.
"""
import numpy as np
import torch
from einops import rearrange
from torch import nn

from .. import kernels as kn


class DualGrainSeperatePermuter(nn.Module):
    """use fine position to represent all-grain position; separate coarse and fine sequences."""

    def __init__(self, coarse_hw=16, fine_hw=32, content_pad_code=1024, content_eos_code=1025,
                 coarse_position_pad_code=256, coarse_position_eos_code=257, fine_position_pad_code=1024,
                 fine_position_eos_code=1025, fine_position_order="region-first") -> None:
        super().__init__()
        self.hw1 = coarse_hw
        self.hw2 = fine_hw // coarse_hw
        self.fine_hw = fine_hw
        self.hw2_square = int(self.hw2 * self.hw2)
        self.content_pad_code = content_pad_code
        self.content_eos_code = content_eos_code
        self.coarse_position_pad_code = coarse_position_pad_code
        self.coarse_position_eos_code = coarse_position_eos_code
        self.fine_position_pad_code = fine_position_pad_code
        self.fine_position_eos_code = fine_position_eos_code
        # plain attributes (not buffers) like the reference: the stage-2 model clones two of them (:69-70)
        self.content_eos_tensor = self.content_eos_code * torch.ones(1).long()
        self.coarse_position_eos_tensor = self.coarse_position_eos_code * torch.ones(1).long()
        self.fine_position_eos_tensor = self.fine_position_eos_code * torch.ones(1).long()
        self.fine_position_order = fine_position_order
        assert self.fine_position_order in ["row-first", "region-first"]
        self.position_sequence_coarse = torch.from_numpy(np.arange(int(coarse_hw ** 2))).long()
        self.position_sequence_fine = torch.from_numpy(np.arange(int(fine_hw ** 2))).long().view(fine_hw, fine_hw)
        if self.fine_position_order == "region-first":
            self.position_sequence_fine = rearrange(self.position_sequence_fine, "(h1 h2) (w1 w2) -> h1 w1 (h2 w2)",
                                                    h1=self.hw1, h2=self.hw2, w1=self.hw1, w2=self.hw2)

    def _codes6(self):
        return (self.content_pad_code, self.content_eos_code, self.coarse_position_pad_code,
                self.coarse_position_eos_code, self.fine_position_pad_code, self.fine_position_eos_code)

    @torch.no_grad()
    def forward(self, indices, grain_indices):
        # grain_indices: 0 for coarse-grained (1 code) and 1 for fine-grained (hw2^2 codes)
        if not indices.is_cuda:
            raise RuntimeError("DualGrainSeperatePermuter (B200) needs CUDA tensors; there is no CPU fallback")
        b = indices.size(0)
        assert indices.shape[1:] == (self.fine_hw, self.fine_hw) and grain_indices.shape[1:] == (self.hw1, self.hw1)
        indices = indices.long().contiguous()
        grain = grain_indices.to(indices.device).long().contiguous()
        flat = grain.view(b, -1)
        # pad_sequence pads to the longest sequence of the batch (+1 for the eos): one 2-int read-back
        n_coarse, n_fine = torch.stack([(flat == 0).sum(1).amax(), (flat == 1).sum(1).amax()]).tolist()
        cc, cp, cs, fc, fp, fs = kn.permuter_forward(
            indices, grain, self.hw1, self.fine_hw, n_coarse + 1, n_fine * self.hw2_square + 1,
            self.fine_position_order == "region-first", self._codes6())
        return {
            "coarse_content": cc, "fine_content": fc,
            "coarse_position": cp, "fine_position": fp,
            "coarse_segment": cs, "fine_segment": fs,
        }

    @torch.no_grad()
    def forward_back(self, coarse_content, fine_content, coarse_position, fine_position):
        if not coarse_content.is_cuda:
            raise RuntimeError("DualGrainSeperatePermuter (B200) needs CUDA tensors; there is no CPU fallback")
        return kn.permuter_backward(coarse_content.long().contiguous(), fine_content.long().contiguous(),
                                    coarse_position.long().contiguous(), fine_position.long().contiguous(),
                                    self.hw1, self.fine_hw, self.coarse_position_eos_code,
                                    self.fine_position_eos_code)
