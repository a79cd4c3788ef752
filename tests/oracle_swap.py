"""Test helper (not a test): swaps the kernel wrappers the `Graph` drop-ins call for the oracle's restatements, so that their host
logic can run on the CPU (the CUDA library has no CPU path).  `install(setattr_fn)` takes pytest's `monkeypatch.setattr` (undone
after the test) or the builtin `setattr` (a spawned worker process that exits afterwards)."""
import torch

from oracle import texpose_oracle as O
from texpose_b200 import camera, ops
from texpose_b200.layers.nerf_static_transient_light import NeRF


class PatchLoss:
    """ops.PatchLoss.apply with the oracle behind it: (losses [render, uncert, trans_reg, all], image_sample, mask_sample)."""

    @staticmethod
    def apply(rgb, uncert, density, image, obj_mask, coords, weights):
        out = O.patch_losses(image, obj_mask, coords, rgb, uncert, density, *weights)
        zero = rgb.sum() * 0
        terms = [out.get(k, zero) for k in ("render", "uncert", "trans_reg")] + [out["all"]]
        return torch.stack(terms), out["image_sample"], out["mask_sample"]


def install(setattr_fn=setattr):
    layers = lambda ml: [(l.weight, l.bias) for l in ml]

    def forward_samples(self, opt, center, ray, depth_samples, latent_variable_trans=None, latent_variable_light=None, mode=None):
        pts = O.points_from_depth(center, ray, depth_samples)
        unit = torch.nn.functional.normalize(ray, dim=-1)[..., None, :].expand_as(pts)
        return O.nerf_stl_forward(pts, unit, latent_variable_trans, latent_variable_light, layers(self.mlp_feat),
                                  layers(self.mlp_rgb), layers(self.mlp_trans))

    def composite(opt, ray, rgb, dens, depth, uncert):
        return O.composite_stl(ray, rgb, dens, depth, uncert, opt.nerf.min_uncert)

    def center_and_ray(opt, pose, intr=None, H=None, W=None, ray_idx=None):
        c, r = O.get_center_and_ray(pose, intr, opt.H, opt.W)
        return (c, r) if ray_idx is None else (O.gather_rays(c, ray_idx), O.gather_rays(r, ray_idx))

    setattr_fn(camera, "get_center_and_ray", center_and_ray)
    setattr_fn(ops, "gather_rows", lambda src, idx: O.gather_rays(src.float(), idx))
    setattr_fn(camera, "view_matrices", lambda pose, intr, one_launch=False: (intr, pose))       # handed through to patch_rays
    setattr_fn(ops, "patch_rays", lambda intr, pose, coords, H, W: O.patch_rays(coords, pose, intr, H, W))
    setattr_fn(ops, "grid_sample_bilinear",
               lambda img, coords: torch.nn.functional.grid_sample(img, coords, mode="bilinear", align_corners=True))
    setattr_fn(ops, "sample_depth",
               lambda zn, zf, N, rand=None, stratified=True, seed=None: O.sample_depth(zn, zf, N, rand if stratified else None))
    setattr_fn(NeRF, "forward_samples", forward_samples)
    setattr_fn(NeRF, "composite", staticmethod(composite))
    setattr_fn(ops, "PatchLoss", PatchLoss)
