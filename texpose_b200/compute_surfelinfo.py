"""Drop-in for the first-party arithmetic of compute_surfelinfo.py (normals from a depth map) and the
depth-guided ("surfel-guided") sampling range of data/lm.py:352-356."""
from __future__ import annotations

from . import camera, ops


def normal_from_depth(pose, depth, intr, h, w, vis=False):
    """compute_surfelinfo.py:37-55 -> [B,3,h,w]."""
    kinv, pinv = camera.view_matrices(pose, intr)
    n = ops.normal_from_depth(kinv, pinv, depth.reshape(-1, h, w))
    if vis:
        n = (n * 0.5 + 0.5) * (depth.reshape(-1, 1, h, w) > 0).float()
    return n


def depth_guided_range(depth, bg_near, bg_far):
    """z_near, z_far = 0.8/1.2 x rendered depth, background range where the depth is 0 (data/lm.py:352-356)."""
    return ops.depth_guided_range(depth, float(bg_near), float(bg_far))
