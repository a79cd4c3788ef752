"""Fused patch gather + ray-wise losses + seeds (csrc/loss.cu) against the reference's own outputs
(tests/golden/loss.npz, generated from Graph.compute_loss + Model.summarize_loss) and the oracle."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import ops, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model.nerf_adapt_st_gan import Graph
from texpose_b200.tools.patch_sampler import FlexPatchSampler

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_patch_loss_matches_reference_golden(golden):
    d = golden("loss")
    rgb, unc, dens = [t.to(DEV).requires_grad_(True) for t in (d.rgb, d.uncert, d.density)]
    w = (float(d.w_render), float(d.w_uncert), float(d.w_trans_reg))
    losses, img_s, mask_s = ops.PatchLoss.apply(rgb, unc, dens, d.image.to(DEV), d.obj_mask.to(DEV), d.coords.to(DEV), w)
    assert torch.equal(img_s.cpu(), d.image_sample)          # bilinear gather: bit-exact
    assert torch.equal(mask_s.cpu(), d.mask_sample)          # nearest gather (align_corners=False): bit-exact
    ref = torch.tensor([float(d.l_render), float(d.l_uncert), float(d.l_trans_reg), float(d.l_all)])
    assert (losses.cpu() - ref).abs().max() <= 1e-5 * ref.abs().max(), (losses.cpu(), ref)
    losses[3].backward()
    assert (rgb.grad.cpu() - d.g_rgb).abs().max() <= 1e-6 * d.g_rgb.abs().max()
    assert (unc.grad.cpu() - d.g_uncert).abs().max() <= 1e-5 * d.g_uncert.abs().max()
    assert (dens.grad.cpu() - d.g_density).abs().max() <= 1e-9


@pytest.mark.parametrize("terms", [(0.0, 0.0, -2.0), (0.0, None, None), (None, 0.5, -1.0)])
def test_patch_loss_vs_oracle_c3_shape(terms):
    """16 patches of 16x16 rays x 128 samples (C3), every subset of terms the yaml can switch off; deterministic."""
    B, P, N, H, W = 16, 16, 128, 128, 128
    g = torch.Generator().manual_seed(5)
    image = torch.rand(B, 3, H, W, generator=g)
    mask = (torch.rand(B, H, W, generator=g) > 0.5).float()
    coords, _ = synth.patch_coords(B, P, seed=4)
    rgb = torch.rand(B, P * P, 3, generator=g)
    unc = torch.rand(B, P * P, 1, generator=g) + 0.05
    dens = torch.rand(B, P * P, N, 2, generator=g) * 4
    r0, u0, d0 = [t.clone().requires_grad_(True) for t in (rgb, unc, dens)]
    ref = O.patch_losses(image, mask, coords, r0, u0, d0, *terms)
    ref["all"].backward()
    r1, u1, d1 = [t.to(DEV).requires_grad_(True) for t in (rgb, unc, dens)]
    args = (image.to(DEV), mask.to(DEV), coords.to(DEV), terms)
    losses, img_s, mask_s = ops.PatchLoss.apply(r1, u1, d1, *args)
    losses[3].backward()
    assert torch.equal(img_s.cpu(), ref["image_sample"]) and torch.equal(mask_s.cpu(), ref["mask_sample"])
    assert abs(losses[3].item() - float(ref["all"])) <= 1e-5 * abs(float(ref["all"]))
    for got, want in ((r1.grad, r0.grad), (u1.grad, u0.grad), (d1.grad, d0.grad)):
        want = want if want is not None else torch.zeros_like(got.cpu())
        assert (got.cpu() - want).abs().max() <= 1e-5 * max(want.abs().max().item(), 1e-12)
    again = ops.PatchLoss.apply(r1.detach(), u1.detach(), d1.detach(), *args)[0]
    assert torch.equal(again, losses.detach())


def test_graph_compute_loss_and_patch_sampler():
    opt = adapt_gan_opt(H=128, W=128, sample_intvs=16, device=DEV)
    opt.b200 = AttrDict(mlp="fp32")
    opt.batch_size, opt.patch_size = 4, 8
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=4).to(DEV)
    # FlexPatchSampler: same three draws as the reference -> identical coords for the same generator state
    torch.manual_seed(77)
    want_c, want_s = O.flex_patch_coords(5, 6)
    torch.manual_seed(77)
    got_c, got_s = FlexPatchSampler()(nbatch=5, patch_size=6, device="cpu")
    assert torch.equal(got_c, want_c) and torch.equal(got_s, want_s)
    var = AttrDict(idx=torch.arange(4, device=DEV))
    var = g.get_ray_idx(opt, var)
    assert var.ray_idx.shape == (4, 8, 8, 2) and var.ray_idx.abs().max() <= 1
    gen = torch.Generator().manual_seed(1)
    var.image = torch.rand(4, 3, 128, 128, generator=gen).to(DEV)
    var.obj_mask = (torch.rand(4, 128, 128, generator=gen) > 0.3).float().to(DEV)
    var.rgb = torch.rand(4, 64, 3, generator=gen).to(DEV).requires_grad_(True)
    var.uncert = (torch.rand(4, 64, 1, generator=gen) + 0.05).to(DEV).requires_grad_(True)
    var.density = torch.rand(4, 64, 16, 2, generator=gen).to(DEV).requires_grad_(True)
    from texpose_b200.model.base import summarize_loss
    loss = summarize_loss(opt, var, g.compute_loss(opt, var, mode="train"))
    ref = O.patch_losses(var.image.cpu(), var.obj_mask.cpu(), var.ray_idx.cpu(), var.rgb.detach().cpu(), var.uncert.detach().cpu(),
                         var.density.detach().cpu(), 0, 0, -2)
    for k in ("render", "uncert", "trans_reg", "all"):
        assert abs(loss[k].item() - float(ref[k])) <= 1e-5 * max(1.0, abs(float(ref[k]))), k
    loss["all"].backward()
    assert var.rgb.grad is not None and var.uncert.grad is not None and var.density.grad is not None


def test_latent_rows_match_embedding_indexing_forward_and_backward():
    """ops.LatentRows == (weight_t[idx], weight_l[idx]) of model/nerf_adapt_st_gan.py:589-603, duplicates included; the table
    gradients equal torch's index backward (sum over the samples of a row) and untouched rows get zeros."""
    from texpose_b200 import ops
    torch.manual_seed(3)
    wt = torch.randn(9, 16, device=DEV, requires_grad=True)
    wl = torch.randn(9, 48, device=DEV, requires_grad=True)
    idx = torch.tensor([3, 0, 3, 7, 7, 7, 1, 0], device=DEV)
    a, b = ops.LatentRows.apply(wt, wl, idx)
    assert torch.equal(a, wt[idx]) and torch.equal(b, wl[idx])
    ga, gb = torch.randn_like(a), torch.randn_like(b)
    (a * ga).sum().backward(retain_graph=True)
    (b * gb).sum().backward()
    ref_t = torch.zeros_like(wt).index_add_(0, idx, ga)
    ref_l = torch.zeros_like(wl).index_add_(0, idx, gb)
    assert (wt.grad - ref_t).abs().max() <= 1e-6 and (wl.grad - ref_l).abs().max() <= 1e-6
    assert wt.grad[2].abs().max() == 0 and wl.grad[8].abs().max() == 0
