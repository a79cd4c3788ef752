"""CPU oracle for the TexPose per-ray NeRF render hot path.

TEST INFRASTRUCTURE ONLY.  This file restates, in plain fp32 PyTorch-on-CPU tensor
arithmetic, the algorithm the reference runs for the path named in BASELINE.json
(`north_star`).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it -- and there only as the checker
or as the timed CPU baseline, never as the product.  The product (`texpose_b200/`)
never imports this module and has no CPU fallback.

Parity status: PINNED.  `oracle/make_golden.py` imports the *real* reference from
/root/reference (stub modules for its missing third-party imports), runs it on the
seeded inputs of `texpose_b200/synth.py` and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against those vectors
(bit-exact for ray/AABB/sample placement, <=2e-6 for the MLP/composite whose fp32
summation order depends on the BLAS build).  The reference itself ships no tests and
no golden vectors, so these generated fixtures are the pin.

All `file:line` citations are relative to /root/reference.
Every function is stateless and device-agnostic (plain torch ops): on CPU tensors it is the oracle / CPU baseline;
`bench.py` also runs it on CUDA tensors to time the reference's eager aten-op path on the same B200 (context only).
Network weights travel as plain lists of (W[out,in], b[out]).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Layer = Tuple[Tensor, Tensor]

# --------------------------------------------------------------------------------------
# poses / rays  (camera.py)
# --------------------------------------------------------------------------------------


def pose_invert(pose: Tensor) -> Tensor:
    """[R|t] -> [R^T | -R^T t]   (camera.py:38-44, Pose.invert with use_inverse=False)."""
    R = pose[..., :3]
    t = pose[..., 3:]
    Rt = R.transpose(-1, -2)
    return torch.cat([Rt, -Rt @ t], dim=-1).float()


def pose_compose_pair(pose_a: Tensor, pose_b: Tensor) -> Tensor:
    """pose_b o pose_a  (camera.py:54-61)."""
    Ra, ta = pose_a[..., :3], pose_a[..., 3:]
    Rb, tb = pose_b[..., :3], pose_b[..., 3:]
    return torch.cat([Rb @ Ra, Rb @ ta + tb], dim=-1).float()


def _append_one(X: Tensor) -> Tensor:
    return torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)  # camera.py:250-253


def unproject_pixels(uv: Tensor, pose: Tensor, intr: Tensor) -> Tuple[Tensor, Tensor]:
    """Pixel coordinates [B,P,2] -> (center[B,P,3], ray[B,P,3]) in the world frame.

    Same operation order as the reference so fp32 cancellation noise matches:
    img2cam = X_h @ K^-1^T (camera.py:266-267), cam2world = X_h @ pose_inv^T
    (camera.py:270-277), ray = grid - center (camera.py:313).
    """
    cam = _append_one(uv) @ intr.inverse().transpose(-1, -2)
    inv_T = pose_invert(pose).transpose(-1, -2)
    grid_w = _append_one(cam) @ inv_T
    center_w = _append_one(torch.zeros_like(cam)) @ inv_T
    return center_w, grid_w - center_w


def get_center_and_ray(pose: Tensor, intr: Tensor, H: int, W: int) -> Tuple[Tensor, Tensor]:
    """Full-frame rays at pixel centres (x+0.5, y+0.5)  (camera.py:292-314)."""
    ys = torch.arange(H, dtype=torch.float32, device=pose.device).add_(0.5)
    xs = torch.arange(W, dtype=torch.float32, device=pose.device).add_(0.5)
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    uv = torch.stack([X, Y], dim=-1).view(-1, 2).repeat(len(pose), 1, 1)
    return unproject_pixels(uv, pose, intr)


def aabb_ray_intersection(aabb_min: Tensor, aabb_max: Tensor, ray_o: Tensor, ray_d: Tensor):
    """Slab test  (camera.py:415-433; duplicate compute_box.py:69-87).  Bit-exact contract."""
    inv_d = torch.reciprocal(ray_d)
    ta = (aabb_min - ray_o) * inv_d
    tb = (aabb_max - ray_o) * inv_d
    t_near = torch.minimum(ta, tb).max(dim=2).values
    t_far = torch.maximum(ta, tb).min(dim=2).values
    valid = (t_far > 0) & (t_far > t_near)
    return t_near, t_far, valid


def enlarge_diagonal(lo: Tensor, hi: Tensor, alpha: float = 0.25):
    """camera.py:436-440."""
    d = hi - lo
    return lo - d * alpha / 2, hi + d * alpha / 2


def padded_box(half_extent: Sequence[float]) -> Tuple[Tensor, Tensor]:
    """Box padding rule of compute_box.py:232-252 for an origin-centred CAD AABB.

    Each corner is pushed along the three axis directions by model.scale/6 (model.scale =
    the largest extent), then the diagonal is enlarged by 25 %.
    """
    h = torch.tensor(half_extent, dtype=torch.float32).view(1, 1, 3)
    pad = (2 * h).max() / 6
    return enlarge_diagonal(-(h + pad), h + pad, alpha=0.25)


def box_bounds_to_range(t_near: Tensor, t_far: Tensor, valid: Tensor, bg_near: float, bg_far: float):
    """compute_box.py:270-271 (zero the misses) + data/lm.py:349-350 (zeros -> background range)."""
    zn = torch.where(valid, t_near, torch.zeros_like(t_near))
    zf = torch.where(valid, t_far, torch.zeros_like(t_far))
    zn = torch.where(zn > 0, zn, torch.full_like(zn, bg_near))
    zf = torch.where(zf > 0, zf, torch.full_like(zf, bg_far))
    return zn, zf


def depth_guided_range(depth: Tensor, bg_near: float, bg_far: float):
    """'render' range source: 0.8/1.2 x depth, zeros -> background  (data/lm.py:352-356)."""
    zn, zf = depth * 0.8, depth * 1.2
    zn = torch.where(zn > 0, zn, torch.full_like(zn, bg_near))
    zf = torch.where(zf > 0, zf, torch.full_like(zf, bg_far))
    return zn, zf


# --------------------------------------------------------------------------------------
# patch rays / bounds  (tools/ray_sampler.py)
# --------------------------------------------------------------------------------------


def patch_rays(coords: Tensor, pose: Tensor, intr: Tensor, H: int, W: int):
    """RaySampler.get_rays (tools/ray_sampler.py:39-69): bilinear lookup of the integer pixel
    ramps (no +0.5) at normalised coords, then the same unprojection as the full frame."""
    B, h, w, _ = coords.shape
    Y, X = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                          indexing="ij")
    X = X[None, None].repeat(B, 1, 1, 1)
    Y = Y[None, None].repeat(B, 1, 1, 1)
    u = F.grid_sample(X, coords, mode="bilinear", align_corners=True)[:, 0]
    v = F.grid_sample(Y, coords, mode="bilinear", align_corners=True)[:, 0]
    uv = torch.stack([u, v], dim=-1).view(B, h * w, 2)
    c, r = unproject_pixels(uv, pose, intr)
    return c.view(B, h, w, 3), r.view(B, h, w, 3)


def patch_bounds(coords: Tensor, z_near: Tensor, z_far: Tensor, H: int, W: int):
    """RaySampler.get_bounds (tools/ray_sampler.py:23-37)."""
    B = coords.shape[0]
    zn = F.grid_sample(z_near.view(B, 1, H, W), coords, mode="bilinear", align_corners=True)[:, 0]
    zf = F.grid_sample(z_far.view(B, 1, H, W), coords, mode="bilinear", align_corners=True)[:, 0]
    return zn, zf


def gather_rays(x: Tensor, ray_idx: Tensor) -> Tensor:
    """Graph.ray_batch_sample (model/nerf_adapt_st_gan.py:702-710) without the hard-coded .cuda()."""
    B, HW, C = x.shape
    flat = ray_idx + HW * torch.arange(B, device=x.device).unsqueeze(1)
    return x.reshape(B * HW, C)[flat].view(B, ray_idx.shape[1], C)


def sample_depth(z_near: Tensor, z_far: Tensor, n: int, rand: Optional[Tensor]) -> Tensor:
    """Stratified depths (model/nerf_adapt_st_gan.py:682-700), metric parametrisation.

    `rand` is the [B,R,N,1] torch.rand draw (or None -> 0.5, sample_stratified false)."""
    lo = z_near[:, :, None, None]
    hi = z_far[:, :, None, None]
    u = rand.clone() if rand is not None else 0.5
    u = u + torch.arange(n, device=z_near.device)[None, None, :, None].float()
    return u / n * (hi - lo) + lo


def points_from_depth(center: Tensor, ray: Tensor, depth: Tensor) -> Tensor:
    """x = c + ray * d for [B,R,N,1] depths (camera.py:317-322, multi_samples=True)."""
    return center[:, :, None] + ray[:, :, None] * depth


# --------------------------------------------------------------------------------------
# the NeRF MLPs  (layers/nerf_static_transient_light.py, layers/nerf.py)
# --------------------------------------------------------------------------------------


def positional_encoding(x: Tensor, L: int) -> Tensor:
    """[...,C] -> [...,2*C*L]; per coordinate [sin f0..f_{L-1}, cos f0..f_{L-1}], f_k = 2^k*pi in fp32
    (layers/nerf_static_transient_light.py:217-223; c2f window off, yaml c2f.range null)."""
    freq = 2 ** torch.arange(L, dtype=torch.float32, device=x.device) * math.pi
    s = x[..., None] * freq
    return torch.stack([s.sin(), s.cos()], dim=-2).flatten(-3)


def _trunk(enc: Tensor, feat_layers: List[Layer], skip: Sequence[int]):
    """8x(Linear+ReLU) with the skip concat; the last layer's row 0 is the raw static density
    (layers/nerf_static_transient_light.py:87-100)."""
    h = enc
    sigma_raw = None
    for li, (Wt, b) in enumerate(feat_layers):
        if li in skip:
            h = torch.cat([h, enc], dim=-1)
        h = F.linear(h, Wt, b)
        if li == len(feat_layers) - 1:
            sigma_raw, h = h[..., 0], h[..., 1:]
        h = F.relu(h)
    return h, sigma_raw


def _head(h: Tensor, layers: List[Layer]) -> Tensor:
    for li, (Wt, b) in enumerate(layers):
        h = F.linear(h, Wt, b)
        if li != len(layers) - 1:
            h = F.relu(h)
    return h


def nerf_stl_forward(points: Tensor, ray_unit: Tensor, latent_trans: Tensor, latent_light: Tensor,
                     feat_layers: List[Layer], rgb_layers: List[Layer], trans_layers: List[Layer],
                     L_3D: int = 10, L_view: int = 4, skip: Sequence[int] = (4,)):
    """Static/transient/light NeRF  (layers/nerf_static_transient_light.py:76-145).

    points, ray_unit: [B,R,N,3]; latents [B,16], [B,48].
    Returns rgb [B,R,N,3,2] (static, transient), density [B,R,N,2], uncert [B,R,N,1]."""
    B, R, N, _ = points.shape
    enc = torch.cat([points, positional_encoding(points, L_3D)], dim=-1)
    with torch.no_grad():
        feat, sigma_raw = _trunk(enc, feat_layers, skip)
        sigma_s = F.softplus(sigma_raw)
    view = torch.cat([ray_unit, positional_encoding(ray_unit, L_view)], dim=-1)
    light = latent_light[:, None, None, :].expand(B, R, N, latent_light.shape[-1])
    rgb_s = torch.sigmoid(_head(torch.cat([feat, view, points, light], dim=-1), rgb_layers))
    trans = latent_trans[:, None, None, :].expand(B, R, N, latent_trans.shape[-1])
    o = _head(torch.cat([feat, trans], dim=-1), trans_layers)
    rgb_t = torch.sigmoid(o[..., :3])
    sigma_t = F.softplus(o[..., 3])
    uncert = F.softplus(o[..., 4:5])
    return torch.stack([rgb_s, rgb_t], dim=-1), torch.stack([sigma_s, sigma_t], dim=-1), uncert


def nerf_plain_forward(points: Tensor, ray_unit: Tensor, feat_layers: List[Layer], rgb_layers: List[Layer],
                       L_3D: int = 10, L_view: int = 4, skip: Sequence[int] = (4,)):
    """Plain NeRF, trunk trainable  (layers/nerf.py:61-99).  Returns rgb [B,R,N,3], density [B,R,N]."""
    enc = torch.cat([points, positional_encoding(points, L_3D)], dim=-1)
    feat, sigma_raw = _trunk(enc, feat_layers, skip)
    view = torch.cat([ray_unit, positional_encoding(ray_unit, L_view)], dim=-1)
    rgb = torch.sigmoid(_head(torch.cat([feat, view, points], dim=-1), rgb_layers))
    return rgb, F.softplus(sigma_raw)


def _intervals(depth: Tensor, ray: Tensor) -> Tensor:
    """delta_i * |ray| with the 1e10 tail  (layers/nerf_static_transient_light.py:169-175)."""
    d = depth[..., 0]
    gap = torch.cat([d[..., 1:] - d[..., :-1], torch.full_like(d[..., :1], 1e10)], dim=2)
    return gap * ray.norm(dim=-1, keepdim=True)


def _transmittance(sd: Tensor) -> Tensor:
    """exp(-exclusive_cumsum(sd))  (layers/nerf_static_transient_light.py:186-191)."""
    shifted = torch.cat([torch.zeros_like(sd[..., :1]), sd[..., :-1]], dim=2)
    return torch.exp(-shifted.cumsum(dim=2))


def composite_stl(ray: Tensor, rgb: Tensor, density: Tensor, depth: Tensor, uncert: Tensor, min_uncert: float):
    """3-chain static/transient/joint volume rendering (layers/nerf_static_transient_light.py:168-212).

    Returns the reference's 11-tuple order: rgb, rgb_static, rgb_transient, depth, opacity,
    opacity_static, opacity_transient, prob, uncert, alpha_static, alpha_transient."""
    dist = _intervals(depth, ray)
    sd_s = density[..., 0] * dist
    sd_t = density[..., 1] * dist
    sd = sd_s + sd_t
    a_s, a_t, a = 1 - torch.exp(-sd_s), 1 - torch.exp(-sd_t), 1 - torch.exp(-sd)
    T, T_s, T_t = _transmittance(sd), _transmittance(sd_s), _transmittance(sd_t)
    p_s, p_t, p = (T * a_s)[..., None], (T * a_t)[..., None], (T * a)[..., None]
    q_s, q_t = (T_s * a_s)[..., None], (T_t * a_t)[..., None]
    out_rgb = (rgb[..., 0] * p_s + rgb[..., 1] * p_t).sum(dim=2)
    out_rgb_s = (q_s * rgb[..., 0]).sum(dim=2)
    out_rgb_t = (q_t * rgb[..., 1]).sum(dim=2)
    out_uncert = (uncert * p_t).sum(dim=2) + min_uncert
    out_depth = (depth * q_s).sum(dim=2)
    return (out_rgb, out_rgb_s, out_rgb_t, out_depth, p.sum(dim=2), q_s.sum(dim=2), q_t.sum(dim=2),
            p, out_uncert, a_s, a_t)


def composite_plain(ray: Tensor, rgb: Tensor, density: Tensor, depth: Tensor,
                    bgcolor: Optional[float] = None):
    """Single-chain compositing  (layers/nerf.py:117-136).  Returns rgb, depth, opacity, prob."""
    sd = density * _intervals(depth, ray)
    p = (_transmittance(sd) * (1 - torch.exp(-sd)))[..., None]
    out_rgb = (rgb * p).sum(dim=2)
    opacity = p.sum(dim=2)
    if bgcolor is not None:
        out_rgb = out_rgb + bgcolor * (1 - opacity)
    return out_rgb, (depth * p).sum(dim=2), opacity, p


def render_stl(center: Tensor, ray: Tensor, z_near: Tensor, z_far: Tensor, rand: Optional[Tensor], n: int,
               latent_trans: Tensor, latent_light: Tensor, feat_layers, rgb_layers, trans_layers,
               min_uncert: float = 0.05, L_3D: int = 10, L_view: int = 4, skip=(4,)):
    """Body of Graph.render after ray selection (model/nerf_adapt_st_gan.py:585-631):
    sample_depth -> forward_samples -> composite -> dict of 11 tensors."""
    depth = sample_depth(z_near, z_far, n, rand)
    pts = points_from_depth(center, ray, depth)
    unit = F.normalize(ray, dim=-1)[..., None, :].expand_as(pts)  # nerf_static_transient_light.py:155-157
    rgb_s, dens, unc = nerf_stl_forward(pts, unit, latent_trans, latent_light, feat_layers, rgb_layers,
                                        trans_layers, L_3D, L_view, skip)
    (rgb, rgb_static, rgb_transient, d, op, op_s, op_t, prob, uncert, a_s, a_t) = composite_stl(
        ray, rgb_s, dens, depth, unc, min_uncert)
    return dict(rgb=rgb, rgb_static=rgb_static, rgb_transient=rgb_transient, opacity=op, opacity_static=op_s,
                opacity_transient=op_t, uncert=uncert, depth=d, alpha_static=a_s, alpha_transient=a_t,
                density=dens)


def nerf_losses(rgb: Tensor, uncert: Tensor, density: Tensor, image: Tensor, mask: Tensor):
    """The three ray-wise loss terms that seed the backward (model/nerf_adapt_st_gan.py:747-763).

    rgb [B,R,3], uncert [B,R,1], density [B,R,N,2]; image [B,R,3], mask [B,R,1] already sampled."""
    render = (mask * ((image - rgb) ** 2 / uncert ** 2)).sum() / (mask.sum() + 1e-5)
    unc = 5 + torch.log(uncert ** 2).mean() / 2
    trans_reg = density[..., -1].mean()
    return render, unc, trans_reg


def sample_patch_targets(image: Tensor, obj_mask: Tensor, coords: Tensor):
    """Patch gather of Graph.compute_loss (model/nerf_adapt_st_gan.py:716-731): image [B,3,H,W] bilinear with
    align_corners=True; mask binarised (>0), sampled with mode='nearest' and the DEFAULT align_corners=False, i.e. pixel
    x = ((c+1)*W-1)/2 rounded half-to-even, zero outside the image.  Returns [B,3,h,w], [B,1,h,w]."""
    B, _, H, W = image.shape
    m = (obj_mask > 0).float().contiguous().view(B, 1, H, W)
    img_s = F.grid_sample(image.contiguous(), coords, mode="bilinear", align_corners=True)
    x = ((coords[..., 0] + 1) * W - 1) / 2
    y = ((coords[..., 1] + 1) * H - 1) / 2
    xi, yi = torch.round(x).long(), torch.round(y).long()          # torch.round = nearbyint (half to even)
    inside = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
    flat = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).view(B, -1)
    vals = torch.gather(m.view(B, H * W), 1, flat).view_as(x)
    mask_s = torch.where(inside, vals, torch.zeros_like(vals))[:, None]
    return img_s, mask_s


def patch_losses(image: Tensor, obj_mask: Tensor, coords: Tensor, rgb: Tensor, uncert: Tensor, density: Tensor,
                 w_render: Optional[float] = 0.0, w_uncert: Optional[float] = 0.0, w_trans_reg: Optional[float] = -2.0):
    """Graph.compute_loss(train_step='nerf') ray-wise terms + Model.summarize_loss (model/nerf_adapt_st_gan.py:712-763,
    model/base.py:145-157): rgb [B,h*w,3], uncert [B,h*w,1], density [B,h*w,N,2]; weights in log10 scale, None = term off."""
    B, h, w, _ = coords.shape
    img_s, mask_s = sample_patch_targets(image, obj_mask, coords)
    rgb_i = rgb.view(B, h, w, 3).permute(0, 3, 1, 2)
    unc_i = uncert.view(B, h, w, 1).permute(0, 3, 1, 2)
    out = dict(image_sample=img_s, mask_sample=mask_s)
    total = 0.0
    if w_render is not None:
        out["render"] = (mask_s * ((img_s - rgb_i) ** 2 / unc_i ** 2)).sum() / (mask_s.sum() + 1e-5)
        total = total + 10 ** float(w_render) * out["render"]
    if w_uncert is not None:
        out["uncert"] = 5 + torch.log(uncert ** 2).mean() / 2
        total = total + 10 ** float(w_uncert) * out["uncert"]
    if w_trans_reg is not None:
        out["trans_reg"] = density[..., -1].mean()
        total = total + 10 ** float(w_trans_reg) * out["trans_reg"]
    out["all"] = total
    return out


def flex_patch_coords(nbatch: int, patch_size: int, min_scale: float = 0.25, max_scale: float = 1.0):
    """FlexPatchSampler.__call__ (tools/patch_sampler.py:80-114) with random_scale = random_shift = True and no annealing:
    three draws from the global torch RNG in the reference's order (scales, h offset, w offset)."""
    lin = torch.linspace(-1, 1, patch_size)
    w, h = torch.meshgrid([lin, lin], indexing="ij")
    h, w = h[None, ..., None], w[None, ..., None]
    scales = torch.rand((nbatch, 1, 1, 1)) * (max_scale - min_scale) + min_scale
    h, w = h * scales, w * scales
    max_offset = 1 - scales
    h_offset = (torch.rand((nbatch, 1, 1, 1)) * 2.0 - 1.0) * max_offset
    w_offset = (torch.rand((nbatch, 1, 1, 1)) * 2.0 - 1.0) * max_offset
    h, w = h + h_offset, w + w_offset
    return torch.cat([h, w], dim=-1).contiguous(), scales.contiguous()


def eval_frame(rgb_static: Tensor, depth: Tensor, image: Tensor, obj_mask: Tensor, H: int, W: int, depth_scale: float):
    """Per-frame part of Model.evaluate_full (model/nerf_adapt_st_gan.py:341-362) for native-resolution frames (the
    interpolate calls at :347-351 are identities then): maps, image * mask, and PSNR = -10 log10 MSE -- per view."""
    B = rgb_static.shape[0]
    rgb_map = rgb_static.view(B, H, W, 3).permute(0, 3, 1, 2)
    depth_map = depth.view(B, H, W, 1).permute(0, 3, 1, 2) / depth_scale
    mask_map = obj_mask.view(B, H, W, 1).permute(0, 3, 1, 2)
    image_masked = image.view(B, 3, H, W) * mask_map
    mse = ((rgb_map - image_masked) ** 2).flatten(1).mean(dim=1)
    return dict(rgb_map=rgb_map, depth_map=depth_map, image_masked=image_masked, mse=mse, psnr=-10 * mse.log10())


# --------------------------------------------------------------------------------------
# surfel info  (compute_surfelinfo.py)
# --------------------------------------------------------------------------------------


def normal_from_depth(pose: Tensor, depth: Tensor, intr: Tensor, H: int, W: int) -> Tensor:
    """Per-pixel normals from a depth map  (compute_surfelinfo.py:37-55)."""
    B = len(depth)
    c, r = get_center_and_ray(pose, intr, H, W)
    P = (c + r * depth.view(B, 1, H * W).permute(0, 2, 1)).permute(0, 2, 1).view(B, 3, H, W)
    tu = P[:, :, 1:-1, 2:] - P[:, :, 1:-1, :-2]
    tv = P[:, :, 2:, 1:-1] - P[:, :, :-2, 1:-1]
    n = F.pad(torch.cross(tu, tv, dim=1), (1, 1, 1, 1))
    n = F.normalize(n, dim=1)
    n[:, -1] *= -1
    return n * (depth[:, None] > 0).float()
