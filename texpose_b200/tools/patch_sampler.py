"""Drop-in for tools/patch_sampler.py `FlexPatchSampler` (reference :63-114): random scale / shift patch coordinates in
[-1, 1].  Host logic: three torch.rand draws in the reference's order (so a seeded run reproduces its patches bit for bit)
and a handful of elementwise ops on a [B,P,P,2] tensor -- no kernel needed."""
from __future__ import annotations

from math import exp

import torch


class FlexPatchSampler:
    def __init__(self, random_shift=True, random_scale=True, min_scale=0.25, max_scale=1.0, scale_anneal=-1):
        self.random_shift = random_shift
        self.random_scale = random_scale
        self.min_scale = min_scale
        self.max_scale = max_scale
        self.scales_curr = (min_scale, max_scale)
        self.iterations = 0
        self.scale_anneal = scale_anneal
        self.full_indices = False

    def __call__(self, nbatch, patch_size, device="cuda"):
        lin = torch.linspace(-1, 1, patch_size, device=device)
        w, h = torch.meshgrid([lin, lin], indexing="ij")
        h, w = h[None, ..., None], w[None, ..., None]
        if self.scale_anneal > 0:
            min_scale = min(0.8, max(self.min_scale, self.max_scale * exp(-self.iterations * self.scale_anneal)))
        else:
            min_scale = self.min_scale
        max_scale = self.max_scale
        self.scales_curr = (min_scale, max_scale)
        if self.random_scale:
            scales = torch.rand((nbatch, 1, 1, 1), device=device) * (max_scale - min_scale) + min_scale
        else:
            scales = torch.ones((nbatch, 1, 1, 1), device=device) * min_scale
        h, w = h * scales, w * scales
        if self.random_shift:
            max_offset = 1 - scales
            h_offset = (torch.rand((nbatch, 1, 1, 1), device=device) * 2.0 - 1.0) * max_offset
            w_offset = (torch.rand((nbatch, 1, 1, 1), device=device) * 2.0 - 1.0) * max_offset
            h, w = h + h_offset, w + w_offset
        return torch.cat([h, w], dim=-1).contiguous(), scales.contiguous()
