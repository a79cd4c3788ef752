"""Drop-in for the arithmetic of compute_box.py: the padded CAD box and the per-pixel AABB sampling bounds.

`box_bounds` reproduces what compute_box.py:266-283 writes to `pred_box_*/NNNNNN.npz` ([2,H,W], misses zeroed);
`box_range` additionally applies data/lm.py:349-350 (zeros -> background range) without the npz round trip.
"""
from __future__ import annotations

import numpy as np
import torch

from . import camera, ops


def padded_box(bb_min, bb_max, scale=None, scale_factor=6, enlarge=0.25):
    """compute_box.py:232-252 for an axis-aligned CAD box: each corner moves outwards along all three axes by
    model.scale / 6 (model.scale defaults to the largest extent), then the diagonal grows by 25 %."""
    bb_min = torch.as_tensor(bb_min, dtype=torch.float32).view(1, 1, 3)
    bb_max = torch.as_tensor(bb_max, dtype=torch.float32).view(1, 1, 3)
    scale = float((bb_max - bb_min).max()) if scale is None else scale
    lo = bb_min - scale / scale_factor
    hi = bb_max + scale / scale_factor
    return camera.enlarge_diagonal(lo, hi, alpha=enlarge)


def get_center_and_ray(pose, intr=None, H=480, W=640):
    """compute_box.py:41-60."""
    kinv, pinv = camera.view_matrices(pose, intr)
    return ops.raygen(kinv, pinv, H, W, 0.5, None)


aabb_ray_intersection = camera.aabb_ray_intersection      # compute_box.py:69-87 duplicates camera.py:415-433


def box_bounds(pose, intr, aabb_min, aabb_max, H=480, W=640):
    """[B,2,H,W] t_near/t_far with misses zeroed (compute_box.py:266-274)."""
    kinv, pinv = camera.view_matrices(pose, intr)
    zn, zf, valid = ops.box_range(kinv, pinv, H, W, aabb_min, aabb_max, 0.0, 0.0, want_valid=True)
    B = zn.shape[0]
    return torch.stack([zn.view(B, H, W), zf.view(B, H, W)], dim=1), valid.view(B, H, W)


def save_box_npz(path, bounds_2hw: torch.Tensor):
    """The reference's on-disk format (compute_box.py:274-283)."""
    np.savez_compressed(path, data=bounds_2hw.detach().cpu().numpy())


def box_range(pose, intr, aabb_min, aabb_max, H, W, bg_near, bg_far):
    """z_near, z_far [B,HW] ready for Graph.nerf_forward (compute_box.py:266-271 + data/lm.py:349-350)."""
    kinv, pinv = camera.view_matrices(pose, intr)
    return ops.box_range(kinv, pinv, H, W, aabb_min, aabb_max, float(bg_near), float(bg_far))
