"""K2 / K2b parity of the fp32 MLP path: forward and the hand-written backward vs the reference fixtures
(tests/golden, generated from /root/reference) and the oracle.  Tolerance 1e-4 max-abs (north_star, fp32)."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200.config import adapt_gan_opt, env_opt
from tests.conftest import check_weight_checksums, layer_list

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _stl(opt):
    from texpose_b200.layers.nerf_static_transient_light import NeRF
    torch.manual_seed(0)
    return NeRF(opt).to(DEV)


def _sub(gr):
    return gr if gr.numel() <= 2048 else gr[:, ::8][::8]


def test_stl_forward_samples_composite_backward_vs_reference(golden):
    g = golden("nerf_stl")
    opt = adapt_gan_opt(device=DEV)
    m = _stl(opt)
    check_weight_checksums(m, g)
    lt = g.latent_trans.to(DEV).requires_grad_(True)
    ll = g.latent_light.to(DEV).requires_grad_(True)
    rgb_s, dens, unc = m.forward_samples(opt, g.center.to(DEV), g.ray.to(DEV), g.depth.to(DEV),
                                         latent_variable_trans=lt, latent_variable_light=ll, mode="train")
    assert rgb_s.shape == (1, 48, 64, 3, 2) and dens.shape == (1, 48, 64, 2) and unc.shape == (1, 48, 64, 1)
    assert (rgb_s.cpu() - g.rgb_samples).abs().max() <= TOL
    assert (dens.cpu() - g.density_samples).abs().max() <= TOL
    assert (unc.cpu() - g.uncert_samples).abs().max() <= TOL
    comp = m.composite(opt, g.ray.to(DEV), rgb_s, dens, g.depth.to(DEV), unc)
    names = ["rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient", "prob",
             "uncert", "alpha_static", "alpha_transient"]
    for k, v in zip(names, comp):
        assert (v.cpu() - g["o_" + k]).abs().max() <= TOL, k
    image, mask = g.image.to(DEV), g.mask.to(DEV)
    rgb, uncert = comp[0], comp[8]
    loss = (mask * ((image - rgb) ** 2 / uncert ** 2)).sum() / (mask.sum() + 1e-5) \
        + (5 + torch.log(uncert ** 2).mean() / 2) + 0.01 * dens[..., -1].mean()
    assert abs(loss.item() - g.loss) <= TOL
    loss.backward()
    assert (lt.grad.cpu() - g.g_latent_trans).abs().max() <= TOL
    assert (ll.grad.cpu() - g.g_latent_light).abs().max() <= TOL
    for name, p in m.named_parameters():
        if name.startswith("mlp_feat") or name == "progress":
            assert p.grad is None, name          # trunk frozen (reference :34,:87)
            continue
        assert p.grad is not None, name
        assert (_sub(p.grad).cpu() - g["g/" + name]).abs().max() <= TOL, name
        assert abs(p.grad.double().sum().item() - g["gsum/" + name]) <= 1e-3, name


def test_stl_forward_explicit_points_matches_forward_samples(golden):
    g = golden("nerf_stl")
    opt = adapt_gan_opt(device=DEV)
    m = _stl(opt)
    pts = O.points_from_depth(g.center, g.ray, g.depth).to(DEV)
    unit = torch.nn.functional.normalize(g.ray, dim=-1).to(DEV)[..., None, :].expand_as(pts)
    with torch.no_grad():
        a = m.forward(opt, pts, ray_unit=unit, latent_variable_trans=g.latent_trans.to(DEV),
                      latent_variable_light=g.latent_light.to(DEV), mode="val")
        b = m.forward(opt, pts, ray_unit=unit.contiguous(), latent_variable_trans=g.latent_trans.to(DEV),
                      latent_variable_light=g.latent_light.to(DEV), mode="val")
    for x, y, ref in zip(a, b, (g.rgb_samples, g.density_samples, g.uncert_samples)):
        assert (x.cpu() - ref).abs().max() <= TOL and (y.cpu() - ref).abs().max() <= TOL
    enc = m.positional_encoding(opt, g.center[0, :5].to(DEV), L=10)
    assert (enc.cpu() - g.posenc_x).abs().max() <= 2e-6


def test_stl_multi_image_batch_vs_oracle():
    """B=3 images with different latents, ragged sizes (S not a multiple of the 128-row tiles)."""
    opt = adapt_gan_opt(device=DEV)
    m = _stl(opt)
    B, R, N = 3, 13, 24
    gen = torch.Generator().manual_seed(4)
    center = torch.randn(B, R, 3, generator=gen) * 0.05 + torch.tensor([0.1, 0.2, -8.0])
    ray = torch.randn(B, R, 3, generator=gen) * 0.1 + torch.tensor([0.0, 0.0, 1.0])
    depth = (torch.rand(B, R, N, 1, generator=gen) + torch.arange(N)[None, None, :, None]) / N * 2 + 7
    lt = torch.randn(B, 16, generator=gen)
    ll = torch.randn(B, 48, generator=gen)
    cpu = _stl(opt).cpu()
    pts = O.points_from_depth(center, ray, depth)
    unit = torch.nn.functional.normalize(ray, dim=-1)[..., None, :].expand_as(pts)
    lt_o, ll_o = lt.clone().requires_grad_(True), ll.clone().requires_grad_(True)
    rl = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in layer_list(cpu.mlp_rgb)]
    tl = [(w.clone().requires_grad_(True), b.clone().requires_grad_(True)) for w, b in layer_list(cpu.mlp_trans)]
    ref = O.nerf_stl_forward(pts, unit, lt_o, ll_o, layer_list(cpu.mlp_feat), rl, tl)
    lt_g, ll_g = lt.to(DEV).requires_grad_(True), ll.to(DEV).requires_grad_(True)
    got = m.forward_samples(opt, center.to(DEV), ray.to(DEV), depth.to(DEV), latent_variable_trans=lt_g,
                            latent_variable_light=ll_g, mode="train")
    cots = [torch.randn(t.shape, generator=gen) / t.numel() * 100 for t in ref]
    for a, b in zip(got, ref):
        assert (a.cpu() - b).abs().max() <= TOL
    torch.autograd.backward(list(ref), cots)
    torch.autograd.backward(list(got), [c.to(DEV) for c in cots])
    assert (lt_g.grad.cpu() - lt_o.grad).abs().max() <= TOL and (ll_g.grad.cpu() - ll_o.grad).abs().max() <= TOL
    for i, (w, b) in enumerate(rl):
        assert (m.mlp_rgb[i].weight.grad.cpu() - w.grad).abs().max() <= TOL, i
        assert (m.mlp_rgb[i].bias.grad.cpu() - b.grad).abs().max() <= TOL, i
    for i, (w, b) in enumerate(tl):
        assert (m.mlp_trans[i].weight.grad.cpu() - w.grad).abs().max() <= TOL, i
        assert (m.mlp_trans[i].bias.grad.cpu() - b.grad).abs().max() <= TOL, i


def test_plain_nerf_forward_backward_vs_reference(golden):
    """layers/nerf.py with the nerf_lm_env.yaml dims: trunk trainable, single-chain composite."""
    g = golden("plain")
    from texpose_b200.layers.nerf import NeRF
    opt = env_opt(device=DEV)
    torch.manual_seed(0)
    m = NeRF(opt).to(DEV)
    check_weight_checksums(m, g)
    rgb_s, dens = m.forward_samples(opt, g.center.to(DEV), g.ray.to(DEV), g.depth.to(DEV), mode="train")
    assert (rgb_s.cpu() - g.rgb_samples).abs().max() <= TOL and (dens.cpu() - g.density_samples).abs().max() <= TOL
    rgb, d, op, prob = m.composite(opt, g.ray.to(DEV), rgb_s, dens, g.depth.to(DEV))
    for a, k in ((rgb, "o_rgb"), (d, "o_depth"), (op, "o_opacity"), (prob, "o_prob")):
        assert (a.cpu() - g[k]).abs().max() <= TOL, k
    loss = ((rgb - g.image.to(DEV)) ** 2).mean() + 0.1 * ((d - 8.0) ** 2).mean() + 0.05 * op.mean()
    assert abs(loss.item() - g.loss) <= TOL
    loss.backward()
    for name, p in m.named_parameters():
        if name == "progress":
            continue
        assert p.grad is not None, name
        assert (_sub(p.grad).cpu() - g["g/" + name]).abs().max() <= TOL, name
