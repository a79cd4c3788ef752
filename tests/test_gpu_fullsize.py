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


def test_c3_backward_is_linear_in_the_upstream_gradient():
    """Size-independent property at the full C3 shape (16 patches x 256 rays x 128 samples): the tensor-core backward is a
    linear map of the loss seeds, and scaling by a power of two is exact in bf16 and fp32 alike -- so doubling the loss must
    double every gradient BIT FOR BIT (dz tiles, thin-gradient MMAs, dW GEMMs, fixed-order reductions)."""
    B, R, N = 16, 256, 128
    opt = adapt_gan_opt(H=128, W=128, sample_intvs=N, device=DEV)
    opt.b200 = AttrDict(mlp="bf16")
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=8).to(DEV)
    gen = torch.Generator().manual_seed(3)
    center = (torch.randn(B, R, 3, generator=gen) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(DEV)
    ray = (torch.randn(B, R, 3, generator=gen) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
    depth = ((torch.rand(B, R, N, 1, generator=gen) + torch.arange(N)[None, None, :, None]) / N * 1.2 + 0.2).to(DEV)
    idx = torch.arange(B, device=DEV) % 8
    target = torch.rand(B, R, 3, generator=gen).to(DEV)
    params = [p for p in g.parameters() if p.requires_grad]

    def grads(scale):
        for p in params:
            p.grad = None
        lt, ll = g.latent_vars_trans.weight[idx], g.latent_vars_light.weight[idx]
        out = g.nerf.forward_samples(opt, center, ray, depth, lt, ll, mode="train")
        comp = g.nerf.composite(opt, ray, *out[:2], depth, out[2])
        loss = ((target - comp[0]) ** 2 / comp[8] ** 2).mean() + torch.log(comp[8] ** 2).mean() + 0.01 * out[1][..., -1].mean()
        (loss * scale).backward()
        return [p.grad.clone() for p in params if p.grad is not None]

    g1, g2 = grads(1.0), grads(2.0)
    assert len(g1) == len(g2) and len(g1) >= 18
    for a, b in zip(g1, g2):
        assert torch.isfinite(a).all() and a.abs().max() > 0
        assert torch.equal(2 * a, b)


def test_c2_full_frame_in_the_parity_mode_on_the_split_kernel():
    """The whole 480 x 640 x 128 frame in the <= 1e-4 mode (split-fp16 tensor-core kernel, 19 launches of 2^21 samples): finite,
    deterministic, row-block shards equal the frame bit for bit, and a strip of object rows equals the SIMT fp32 kernels to 1e-4."""
    from texpose_b200 import _C
    H, W, N = 480, 640, 128
    opt, g, pose, intr, zn, zf = _setup(H, W, N, "fp32")
    dr = (zn[:, :, None], zf[:, :, None])
    _C.launch_counts.clear()
    with torch.no_grad():
        full = g.nerf_forward(opt, AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=torch.ones(1, H, W, device=DEV),
                                            idx=torch.zeros(1, dtype=torch.long, device=DEV)), mode="val")
    assert _C.launch_counts.get("tp_tc32_forward", 0) >= 1 and "tp_linear_forward" not in _C.launch_counts
    for k in ("rgb", "depth", "uncert", "opacity"):
        assert torch.isfinite(full[k]).all(), k
    assert (full.opacity - 1).abs().max() < 1e-4
    b, e = parallel.shard_rays(H * W, 3, 8, align=W)
    idx = torch.arange(b, e, device=DEV)[None]
    opt_s = AttrDict(opt)
    opt_s.b200 = AttrDict(mlp="fp32", fp32_engine="simt")
    with torch.no_grad():
        part = g.render(opt, pose, intr=intr, ray_idx=idx, depth_range=dr, mode="val")
        part2 = g.render(opt, pose, intr=intr, ray_idx=idx, depth_range=dr, mode="val")
        strip = torch.arange(b + 60 * W, b + 60 * W + 2048, device=DEV)[None]        # rows through the object
        tc = g.render(opt, pose, intr=intr, ray_idx=strip, depth_range=dr, mode="val")
        simt = g.render(opt_s, pose, intr=intr, ray_idx=strip, depth_range=dr, mode="val")
    for k in ("rgb", "depth", "uncert", "opacity", "rgb_static"):
        assert torch.equal(part[k], part2[k]) and torch.equal(part[k], full[k][:, b:e]), k
        assert (tc[k] - simt[k]).abs().max() <= 1e-4, (k, float((tc[k] - simt[k]).abs().max()))
