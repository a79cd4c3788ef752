"""Size-independent properties at BASELINE.json's full sizes (C2: 480x640 rays x 128 samples, bf16 tensor-core path):
ray shards rendered separately are bit-identical to the whole frame (the multi-GPU partition is by ray, same kernel, same
per-ray arithmetic), renders are deterministic, opacity == 1 with the 1e10 tail, weights sum to opacity."""
import pytest
import torch

from texpose_b200 import compute_box, parallel, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model.nerf_adapt_st_gan import Graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(H, W, N, precision):
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.nerf.sample_stratified = False
    opt.b200 = AttrDict(mlp=precision)
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=2).to(DEV).eval()
    pose, intr = synth.poses([0]).to(DEV), synth.intrinsics(1).to(DEV)
    intr = intr.clone()
    intr[:, :2] *= W / 640
    lo, hi = [t.to(DEV) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
    return opt, g, pose, intr, zn, zf


def test_c2_full_frame_properties_and_shard_equality():
    H, W, N = 480, 640, 128
    opt, g, pose, intr, zn, zf = _setup(H, W, N, "bf16")
    dr = (zn[:, :, None], zf[:, :, None])
    with torch.no_grad():
        idx = torch.arange(H * W, device=DEV)[None]
        full = g.render(opt, pose, intr=intr, ray_idx=idx, depth_range=dr, mode="val")
        again = g.render(opt, pose, intr=intr, ray_idx=idx, depth_range=dr, mode="val")
    assert full.rgb.shape == (1, H * W, 3) and full.density.shape == (1, H * W, N, 2)
    for k in ("rgb", "depth", "uncert", "opacity"):
        assert torch.equal(full[k], again[k]), k                         # deterministic
        assert torch.isfinite(full[k]).all(), k
    assert (full.opacity - 1).abs().max() < 1e-4 and (full.opacity_static - 1).abs().max() < 1e-4
    # own-chain composites are convex combinations; `rgb` is not (both terms use the joint transmittance and
    # alpha_s + alpha_t >= alpha: layers/nerf_static_transient_light.py:193-205), so it is only bounded by 2
    assert (full.rgb_static >= 0).all() and (full.rgb_static <= 1 + 1e-4).all()
    assert (full.rgb_transient >= 0).all() and (full.rgb_transient <= 1 + 1e-4).all()
    assert (full.rgb >= 0).all() and (full.rgb <= 2 + 1e-4).all()
    assert (full.depth >= zn[:, :, None] - 1e-3).all() and (full.depth <= zf[:, :, None] + 1e-3).all()
    # 8-way row-block ray shards (what 8 GPUs would each render) == the whole frame, bit for bit
    world = 8
    for rank in (0, 3, 7):
        b, e = parallel.shard_rays(H * W, rank, world, align=W)
        with torch.no_grad():
            part = g.render(opt, pose, intr=intr, ray_idx=torch.arange(b, e, device=DEV)[None], depth_range=dr, mode="val")
        for k in ("rgb", "depth", "uncert", "opacity", "rgb_static"):
            assert torch.equal(part[k], full[k][:, b:e]), (rank, k)


def test_fp32_and_bf16_agree_on_a_frame_strip():
    """The two arithmetic modes on the same 16 image rows of the 480x640 frame (object rows)."""
    H, W, N = 480, 640, 64
    opt32, g, pose, intr, zn, zf = _setup(H, W, N, "fp32")
    opt16 = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt16.nerf.sample_stratified = False
    opt16.b200 = AttrDict(mlp="bf16")
    dr = (zn[:, :, None], zf[:, :, None])
    idx = torch.arange(232 * W, 248 * W, device=DEV)[None]
    with torch.no_grad():
        a = g.render(opt32, pose, intr=intr, ray_idx=idx, depth_range=dr, mode="val")
        b = g.render(opt16, pose, intr=intr, ray_idx=idx, depth_range=dr, mode="val")
    box = (zf[0, idx[0]] < 29)                                            # AABB-bounded rays (the parity contract)
    assert box.sum() > 500
    assert (a.rgb - b.rgb)[:, box].abs().max() <= 1e-2
    assert (a.depth - b.depth)[:, box].abs().max() <= 1e-2
    assert (a.opacity - b.opacity).abs().max() <= 1e-2
    assert ((a.depth - b.depth)[:, ~box].abs() / 30).max() <= 1e-2       # background rays sample depths in [0, 30]
