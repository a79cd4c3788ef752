"""K3 parity: compositing forward and backward vs the oracle (autograd on CPU), fp32 tolerance 1e-4
(north_star), through the C-ABI.  Covers ragged N (not a multiple of 32), N=1, zero densities."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import ops
from texpose_b200.config import adapt_gan_opt
from texpose_b200.layers.nerf_static_transient_light import NeRF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ["rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient", "prob",
         "uncert", "alpha_static", "alpha_transient"]


def _inputs(B, R, N, seed, zero_density=False):
    g = torch.Generator().manual_seed(seed)
    ray = torch.randn(B, R, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 1.0])
    rgb = torch.rand(B, R, N, 3, 2, generator=g)
    den = torch.rand(B, R, N, 2, generator=g) * 3
    if zero_density:
        den[:, ::2] = 0
    depth = (torch.rand(B, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 2.5 + 6.7
    unc = torch.rand(B, R, N, 1, generator=g) + 0.01
    return ray, rgb, den, depth, unc


@pytest.mark.parametrize("B,R,N,zero", [(1, 33, 64, False), (2, 17, 128, False), (1, 9, 48, False), (1, 5, 1, False),
                                        (1, 7, 200, True)])
def test_composite_stl_forward_backward(B, R, N, zero):
    ray, rgb, den, depth, unc = _inputs(B, R, N, 10 + N, zero)
    rgb_o, den_o, unc_o = (t.clone().requires_grad_(True) for t in (rgb, den, unc))
    ref = O.composite_stl(ray, rgb_o, den_o, depth, unc_o, 0.05)
    rgb_g, den_g, unc_g = (t.to(DEV).requires_grad_(True) for t in (rgb, den, unc))
    got = NeRF.composite(adapt_gan_opt(device=DEV), ray.to(DEV), rgb_g, den_g, depth.to(DEV), unc_g)
    assert len(got) == 11
    for k, a, b in zip(NAMES, got, ref):
        assert a.shape == b.shape, k
        assert (a.cpu() - b).abs().max() <= 1e-4, (k, (a.cpu() - b).abs().max())
    # seeds on every returned tensor (API generality), fixed random cotangents
    gen = torch.Generator().manual_seed(5)
    cots = [torch.randn(t.shape, generator=gen) for t in ref]
    torch.autograd.backward(list(ref), cots)
    torch.autograd.backward(list(got), [c.to(DEV) for c in cots])
    for name, a, b in (("rgb", rgb_g, rgb_o), ("density", den_g, den_o), ("uncert", unc_g, unc_o)):
        scale = max(1.0, b.grad.abs().max().item())
        err = (a.grad.cpu() - b.grad).abs().max().item()
        assert err <= 1e-4 * scale, (name, err, scale)


def test_composite_stl_training_seeds_only(golden):
    """Only rgb / uncert receive gradients in training (model/nerf_adapt_st_gan.py:747-763)."""
    g = golden("nerf_stl")
    rgb_o = g.rgb_samples.clone().requires_grad_(True)
    den_o = g.density_samples.clone().requires_grad_(True)
    unc_o = g.uncert_samples.clone().requires_grad_(True)
    ref = O.composite_stl(g.ray, rgb_o, den_o, g.depth, unc_o, 0.05)
    l = O.nerf_losses(ref[0], ref[8], den_o, g.image, g.mask)
    (l[0] + l[1] + 0.01 * l[2]).backward()
    rgb_g, den_g, unc_g = (t.to(DEV).requires_grad_(True) for t in (g.rgb_samples, g.density_samples, g.uncert_samples))
    got = ops.CompositeSTL.apply(g.ray.to(DEV), rgb_g, den_g, g.depth.to(DEV), unc_g, 0.05)
    for k, a in zip(NAMES, got):
        assert (a.cpu() - g["o_" + k]).abs().max() <= 1e-4, k
    image, mask = g.image.to(DEV), g.mask.to(DEV)
    loss = (mask * ((image - got[0]) ** 2 / got[8] ** 2)).sum() / (mask.sum() + 1e-5) \
        + (5 + torch.log(got[8] ** 2).mean() / 2) + 0.01 * den_g[..., -1].mean()
    assert abs(loss.item() - g.loss) <= 1e-4
    loss.backward()
    for a, b in ((rgb_g, rgb_o), (den_g, den_o), (unc_g, unc_o)):
        assert (a.grad.cpu() - b.grad).abs().max() <= 1e-4


@pytest.mark.parametrize("N,bg", [(64, None), (40, 0.7)])
def test_composite_plain(N, bg):
    ray, rgb, den, depth, _ = _inputs(2, 11, N, 3)
    rgb, den = rgb[..., 0].contiguous(), den[..., 0].contiguous()
    rgb_o, den_o = rgb.clone().requires_grad_(True), den.clone().requires_grad_(True)
    ref = O.composite_plain(ray, rgb_o, den_o, depth, bg)
    rgb_g, den_g = rgb.to(DEV).requires_grad_(True), den.to(DEV).requires_grad_(True)
    got = ops.CompositePlain.apply(ray.to(DEV), rgb_g, den_g, depth.to(DEV), bg)
    for a, b in zip(got, ref):
        assert (a.cpu() - b).abs().max() <= 1e-4
    gen = torch.Generator().manual_seed(6)
    cots = [torch.randn(t.shape, generator=gen) for t in ref]
    torch.autograd.backward(list(ref), cots)
    torch.autograd.backward(list(got), [c.to(DEV) for c in cots])
    for a, b in ((rgb_g, rgb_o), (den_g, den_o)):
        scale = max(1.0, b.grad.abs().max().item())
        assert (a.grad.cpu() - b.grad).abs().max() <= 1e-4 * scale


def test_full_size_properties():
    """C2-sized property checks that need no oracle: opacity == 1 - prod(1-alpha) -> ~1 with the 1e10 tail,
    weights sum to opacity, static chain independent of the transient density."""
    R, N = 480 * 640 // 8, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    ray = torch.randn(1, R, 3, device=DEV, generator=g) * 0.1 + torch.tensor([0, 0, 1.0], device=DEV)
    rgb = torch.rand(1, R, N, 3, 2, device=DEV, generator=g)
    den = torch.rand(1, R, N, 2, device=DEV, generator=g)
    depth = ((torch.rand(1, R, N, 1, device=DEV, generator=g) + torch.arange(N, device=DEV)[None, None, :, None]) / N) * 2 + 7
    unc = torch.rand(1, R, N, 1, device=DEV, generator=g)
    out = ops.CompositeSTL.apply(ray, rgb, den, depth, unc, 0.05)
    assert (out[4] - 1).abs().max() < 1e-4 and (out[5] - 1).abs().max() < 1e-4 and (out[6] - 1).abs().max() < 1e-4
    assert (out[7].sum(dim=2) - out[4]).abs().max() < 1e-5
    den2 = den.clone(); den2[..., 1] *= 3
    out2 = ops.CompositeSTL.apply(ray, rgb, den2, depth, unc, 0.05)
    assert torch.equal(out2[1], out[1]) and torch.equal(out2[3], out[3]) and torch.equal(out2[9], out[9])
    assert (out[0] >= 0).all() and (out[0] <= 1 + 1e-4).all()


@pytest.mark.parametrize("N", [64, 128])
def test_composite_stl_vectorised_matches_generic(N):
    """N = 64 / 128 take the lane-owns-K-consecutive-samples kernel; a misaligned `prob` buffer forces the generic
    warp-chunk kernel through the same entry point.  Same math, different summation order: agreement to fp32 round-off,
    including saturated (huge density), empty (zero density) and NaN-free tails."""
    from texpose_b200 import _C
    B, R = 2, 301
    ray, rgb, den, depth, unc = _inputs(B, R, N, 77)
    den[0, :40] = 0.0                 # empty rays
    den[1, :40] *= 1e4                # saturate within the first samples
    den[1, 40:80, N // 2:] = 0.0      # density only in the first half
    t = [x.to(DEV).contiguous() for x in (ray, rgb, den, depth, unc)]
    outs = {}
    for name, off in (("vec", 0), ("generic", 1)):
        o3 = [torch.empty(B * R * 3, device=DEV) for _ in range(3)]
        o1 = [torch.empty(B * R, device=DEV) for _ in range(5)]
        prob_buf = torch.empty(B * R * N + 4, device=DEV)
        prob = prob_buf[off:off + B * R * N]
        a_s, a_t = torch.empty(B * R * N, device=DEV), torch.empty(B * R * N, device=DEV)
        _C.call("tp_composite_stl_forward", *[ops._p(x) for x in t], B * R, N, 0.05, ops._p(o3[0]), ops._p(o3[1]),
                ops._p(o3[2]), ops._p(o1[0]), ops._p(o1[1]), ops._p(o1[2]), ops._p(o1[3]), prob.data_ptr(), ops._p(o1[4]),
                ops._p(a_s), ops._p(a_t), ops._stream())
        torch.cuda.synchronize()
        outs[name] = [x.clone() for x in (*o3, *o1, prob, a_s, a_t)]
    for i, (a, b) in enumerate(zip(outs["vec"], outs["generic"])):
        assert torch.isfinite(a).all(), i
        assert (a - b).abs().max() <= 2e-6 * max(1.0, b.abs().max().item()), (i, (a - b).abs().max())
