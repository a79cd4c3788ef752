"""Seeded LineMOD-shaped synthetic inputs.

No dataset is available offline, so every test/bench input is generated here from a seed:
the LineMOD camera matrix (compute_box.py:166-168), a random rotation with the object 0.8 m
in front of the camera (metres x depth.scale 10, data/lm.py:385,408), and a duck-like CAD box.
Host-side input generation only -- no render arithmetic lives in this file.
"""
from __future__ import annotations

import torch

LM_K = [[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]]
LM_H, LM_W = 480, 640
DUCK_HALF_EXTENT = [0.522, 0.387, 0.428]     # stand-in CAD half extents, units m x 10
OBJ_T = [0.3, -0.2, 8.0]
BG_RANGE = (0.0, 30.0)                        # nerf.depth.range [0,3] x scale 10


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    """Unit quaternion (a,b,c,d) -> rotation matrix (standard Hamilton convention)."""
    a, b, c, d = q.unbind(-1)
    rows = [
        torch.stack([1 - 2 * (c * c + d * d), 2 * (b * c - a * d), 2 * (a * c + b * d)], -1),
        torch.stack([2 * (b * c + a * d), 1 - 2 * (b * b + d * d), 2 * (c * d - a * b)], -1),
        torch.stack([2 * (b * d - a * c), 2 * (a * b + c * d), 1 - 2 * (b * b + c * c)], -1),
    ]
    return torch.stack(rows, -2)


def intrinsics(B: int = 1) -> torch.Tensor:
    return torch.tensor(LM_K, dtype=torch.float32).repeat(B, 1, 1)


def pose_for_seed(seed: int) -> torch.Tensor:
    """[3,4] object->camera pose: R from a normalised randn(4) quaternion, t = OBJ_T."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    R = quat_to_rot(q)
    t = torch.tensor(OBJ_T, dtype=torch.float32)
    return torch.cat([R, t[:, None]], dim=-1).float()


def poses(seeds) -> torch.Tensor:
    return torch.stack([pose_for_seed(s) for s in seeds], 0)


def padded_aabb(half_extent=DUCK_HALF_EXTENT):
    """Box rule of compute_box.py:232-252 for an origin-centred box: every side grows by
    max-extent/6, then the diagonal is enlarged by 25 % (alpha/2 per side)."""
    h = torch.tensor(half_extent, dtype=torch.float32)
    h = h + (2 * h).max() / 6
    lo, hi = -h, h
    d = hi - lo
    return (lo - d * 0.25 / 2).view(1, 1, 3), (hi + d * 0.25 / 2).view(1, 1, 3)


def latents(B: int, seed: int = 1, n_trans: int = 16, n_light: int = 48):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, n_trans, generator=g), torch.randn(B, n_light, generator=g)


def patch_coords(B: int, P: int, seed: int = 2, min_scale: float = 0.25, max_scale: float = 1.0):
    """FlexPatchSampler-shaped coords [B,P,P,2] in [-1,1] (tools/patch_sampler.py:80-114):
    a regular PxP lattice scaled by s~U[min,max] and shifted by at most 1-s."""
    g = torch.Generator().manual_seed(seed)
    lin = torch.linspace(-1, 1, P)
    w, h = torch.meshgrid(lin, lin, indexing="ij")
    s = torch.rand(B, 1, 1, 1, generator=g) * (max_scale - min_scale) + min_scale
    off_h = (torch.rand(B, 1, 1, 1, generator=g) * 2 - 1) * (1 - s)
    off_w = (torch.rand(B, 1, 1, 1, generator=g) * 2 - 1) * (1 - s)
    hh = h[None, ..., None] * s + off_h
    ww = w[None, ..., None] * s + off_w
    return torch.cat([hh, ww], dim=-1).contiguous(), s


def ellipsoid_depth(pose: torch.Tensor, intr: torch.Tensor, H: int, W: int, radii=DUCK_HALF_EXTENT):
    """Analytic camera-z depth of an origin-centred ellipsoid (0 where missed) -- the synthetic
    stand-in for the rasterised depth compute_surfelinfo.py:114-115 would provide."""
    ys = torch.arange(H, dtype=torch.float64) + 0.5
    xs = torch.arange(W, dtype=torch.float64) + 0.5
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    Kinv = torch.linalg.inv(intr.double())
    pix = torch.stack([X, Y, torch.ones_like(X)], -1).view(-1, 3)
    out = []
    for b in range(len(pose)):
        R, t = pose[b, :, :3].double(), pose[b, :, 3].double()
        dirs = (pix @ Kinv[b].T) @ R          # world-frame ray with camera-z == 1
        o = -(R.T @ t)
        r = torch.tensor(radii, dtype=torch.float64)
        dn, on = dirs / r, o / r
        a = (dn * dn).sum(-1)
        bq = 2 * (dn * on).sum(-1)
        c = (on * on).sum() - 1
        disc = bq * bq - 4 * a * c
        tt = (-bq - disc.clamp(min=0).sqrt()) / (2 * a)
        out.append(torch.where((disc > 0) & (tt > 0), tt, torch.zeros_like(tt)).view(H, W))
    return torch.stack(out, 0).float()


def icosphere(subdiv=2, radius=1.0):
    """Synthetic CAD model for the rasteriser: icosphere vertices [V,3] float32 and faces [F,3] int32 (20 x 4^subdiv)."""
    import numpy as np
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1),
         (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []
        for a, b, c in f:
            m = []
            for i, j in ((a, b), (b, c), (c, a)):
                k = (min(i, j), max(i, j))
                if k not in cache:
                    p = v[i] + v[j]
                    v.append(p / np.linalg.norm(p))
                    cache[k] = len(v) - 1
                m.append(cache[k])
            nf += [(a, m[0], m[2]), (b, m[1], m[0]), (c, m[2], m[1]), (m[0], m[1], m[2])]
        f = nf
    return torch.from_numpy((np.array(v) * radius).astype(np.float32)), torch.from_numpy(np.array(f, np.int32))
