"""Edge cases through the drop-in surface: empty ray sets (an eval frame whose mask prior selects nothing, a rank whose
shard is empty), single-ray / single-sample shapes and ragged sample counts, in both MLP modes.  The reference's torch ops
return empty tensors of the right shape for empty inputs; so must the kernels' wrappers (no launch, no error)."""
import pytest
import torch

from texpose_b200 import camera, ops, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.layers.nerf_static_transient_light import NeRF
from texpose_b200.model.nerf_adapt_st_gan import Graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _opt(mlp, N=64, H=48, W=64):
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.b200 = AttrDict(mlp=mlp)
    return opt


@pytest.mark.parametrize("mlp", ["fp32", "bf16"])
def test_empty_ray_set(mlp):
    opt = _opt(mlp)
    torch.manual_seed(0)
    m = NeRF(opt).to(DEV)
    pose, intr = synth.poses([0]).to(DEV), synth.intrinsics(1).to(DEV)
    idx = torch.zeros(1, 0, dtype=torch.long, device=DEV)
    center, ray = camera.get_center_and_ray(opt, pose, intr=intr, ray_idx=idx)
    assert center.shape == (1, 0, 3) and ray.shape == (1, 0, 3)
    lo, hi = [t.to(DEV) for t in synth.padded_aabb()]
    tn, tf, valid = camera.aabb_ray_intersection(lo.view(1, 1, 3), hi.view(1, 1, 3), center, ray)
    assert tn.shape == (1, 0) and valid.dtype == torch.bool
    depth = ops.sample_depth(tn, tf, opt.nerf.sample_intvs, rand=torch.rand(1, 0, opt.nerf.sample_intvs, 1, device=DEV))
    assert depth.shape == (1, 0, opt.nerf.sample_intvs, 1)
    lt, ll = [t.to(DEV) for t in synth.latents(1)]
    with torch.no_grad():
        rgb, den, unc = m.forward_samples(opt, center, ray, depth, lt, ll, mode="val")
        out = m.composite(opt, ray, rgb, den, depth, unc)
    assert rgb.shape == (1, 0, opt.nerf.sample_intvs, 3, 2) and den.shape == (1, 0, opt.nerf.sample_intvs, 2)
    assert len(out) == 11 and out[0].shape == (1, 0, 3) and out[7].shape == (1, 0, opt.nerf.sample_intvs, 1)
    torch.cuda.synchronize()


@pytest.mark.parametrize("mlp", ["fp32", "bf16"])
@pytest.mark.parametrize("R,N", [(1, 1), (1, 64), (3, 37), (130, 2)])
def test_tiny_and_ragged_shapes_agree_between_modes(mlp, R, N):
    """One ray, one sample, N not a multiple of 32 / 4, a tile boundary inside a ray: finite outputs of the right shape, and
    the bf16 tensor-core path stays within its 1e-2 contract of the fp32 kernels on the rendered values."""
    opt = _opt(mlp, N=N)
    torch.manual_seed(0)
    m = NeRF(opt).to(DEV)
    g = torch.Generator().manual_seed(R * 1000 + N)
    center = (torch.randn(1, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -8.0])).to(DEV)
    ray = (torch.randn(1, R, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
    depth = ((torch.rand(1, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 2.5 + 6.7).to(DEV)
    lt, ll = [t.to(DEV) for t in synth.latents(1)]
    with torch.no_grad():
        rgb, den, unc = m.forward_samples(opt, center, ray, depth, lt, ll, mode="val")
        out = m.composite(opt, ray, rgb, den, depth, unc)
        opt32 = _opt("fp32", N=N)
        rgb32, den32, unc32 = m.forward_samples(opt32, center, ray, depth, lt, ll, mode="val")
        out32 = m.composite(opt32, ray, rgb32, den32, depth, unc32)
    assert rgb.shape == (1, R, N, 3, 2) and den.shape == (1, R, N, 2) and unc.shape == (1, R, N, 1)
    for a, b in zip(out, out32):
        assert a.shape == b.shape and torch.isfinite(a).all()
    tol = 1e-2 if mlp == "bf16" else 0.0
    for k in (0, 3, 4):      # rgb, depth, opacity
        assert (out[k] - out32[k]).abs().max() <= tol, (k, (out[k] - out32[k]).abs().max())


def test_eval_frame_with_empty_mask_keeps_defaults():
    """Graph.render_by_slices(mode='eval') with a mask prior that selects no pixel: the defaults of
    model/nerf_adapt_st_gan.py:657-667 come back untouched and nothing is launched on an empty ray set."""
    opt = _opt("bf16", N=64)
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=4).to(DEV)
    g.eval()
    pose, intr = synth.poses([0]).to(DEV), synth.intrinsics(1).to(DEV)
    HW = opt.H * opt.W
    zn, zf = torch.full((1, HW), 6.0, device=DEV), torch.full((1, HW), 9.0, device=DEV)
    mask = torch.zeros(1, opt.H, opt.W, device=DEV)
    with torch.no_grad():
        ret = g.render_by_slices(opt, pose, intr=intr, depth_range=(zn[:, :, None], zf[:, :, None]), object_mask=mask,
                                 sample_idx=torch.tensor(0, device=DEV), mode="eval")
    assert ret.rgb.shape == (1, HW, 3) and ret.opacity.shape == (1, HW, 1)
    assert torch.isfinite(ret.rgb).all() and float(ret.opacity.abs().max()) == 0.0
