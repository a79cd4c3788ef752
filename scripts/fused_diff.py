"""Per-output max-abs difference between the fused render launch and the multi-kernel path (same rays, depths, weights).
Debugging aid.  usage: [TEXPOSE_B200_LIB=...] python scripts/fused_diff.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import compute_box, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402

dev = "cuda:0"
KEYS = ("rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient", "uncert",
        "alpha_static", "alpha_transient", "density")
for N in (128, 64, 32):
    H, W = 24, 40
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=dev)
    opt.nerf.sample_stratified = False
    opt.b200 = AttrDict(mlp="bf16")
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=4).to(dev).eval()
    pose, intr = synth.poses([0, 3]).to(dev), synth.intrinsics(2).to(dev).clone()
    intr[:, :2] *= 0.06
    lo, hi = [t.to(dev) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
    dr = (zn[:, :, None], zf[:, :, None])
    o2 = AttrDict(opt); o2.b200 = AttrDict(opt.b200); o2.b200.fused_render = False
    gen = torch.Generator().manual_seed(5)
    idx = torch.stack([torch.randperm(H * W, generator=gen)[:333] for _ in range(2)]).to(dev)
    for name, ray_idx in (("random", idx), ("block", range(80, 80 + 7 * W + 3))):
        with torch.no_grad():
            a = g.render(opt, pose, intr=intr, ray_idx=ray_idx, depth_range=dr, mode="val")
            a2 = g.render(opt, pose, intr=intr, ray_idx=ray_idx, depth_range=dr, mode="val")
            b = g.render(o2, pose, intr=intr, ray_idx=ray_idx, depth_range=dr, mode="val")
        print(f"N={N} {name}: " + " ".join(f"{k}={float((a[k] - b[k]).abs().max()):.1e}" for k in KEYS),
              "| repeat-identical:", all(torch.equal(a[k], a2[k]) for k in KEYS))
