"""Drop-in behaviour the reference engine relies on (model/nerf_adapt_st_gan.py:100-127, model/base.py:145-157):
toggle_grad(nerf, True) before every step must not un-freeze the static trunk, val renders broadcast latent row 0 to every
view, and every term of compute_loss is differentiable on its own (summarize_loss of the reference builds `all`)."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import _C, compute_box, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model import base
from texpose_b200.model.nerf_adapt_st_gan import Graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _train_inputs(B, P, H, W):
    pose = synth.poses(list(range(B))).to(DEV)
    K = torch.tensor([[572.4114, 0, W / 2 - 572.4114 * 0.3 / 8], [0, 573.57043, H / 2 + 573.57043 * 0.2 / 8], [0, 0, 1]])
    intr = K.repeat(B, 1, 1).to(DEV)
    lo, hi = [t.to(DEV) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
    coords, _ = synth.patch_coords(B, P, seed=2)
    g = torch.Generator().manual_seed(8)
    image = torch.rand(B, 3, H, W, generator=g).to(DEV)
    mask = (torch.rand(B, H, W, generator=g) > 0.3).float().to(DEV)
    return pose, intr, zn, zf, coords.to(DEV), image, mask


def test_toggle_grad_keeps_the_static_trunk_frozen_and_on_tensor_cores():
    """the engine sets requires_grad=True on every nerf parameter before each nerf_trainstep (:110); the
    reference's trunk runs under torch.no_grad() (layers/nerf_static_transient_light.py:87-101) so it still gets no gradient.
    Here: trunk grads stay None, head grads arrive, and the step runs the fused tcgen05 launches (not the fp32 SIMT path)."""
    B, P, H, W, N = 4, 16, 128, 128, 64
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.b200 = AttrDict(mlp="auto")
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=4).to(DEV)
    for p in g.nerf.parameters():                     # Model.toggle_grad(self.graph.nerf, True)
        p.requires_grad_(True)
    assert g.nerf.uses_tensor_cores(opt, "train")
    pose, intr, zn, zf, coords, image, mask = _train_inputs(B, P, H, W)
    idx = torch.arange(B, device=DEV)
    _C.launch_counts.clear()
    ret = g.render(opt, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]), sample_idx=idx, mode="train")
    var = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
    var.update(ret)
    loss = base.summarize_loss(opt, var, g.compute_loss(opt, var, mode="train"))
    loss["all"].backward()
    assert _C.launch_counts.get("tp_tc_nerf_stl_forward", 0) == 1 and _C.launch_counts.get("tp_tc_heads_backward", 0) == 1
    assert _C.launch_counts.get("tp_linear_forward", 0) == 0 and _C.launch_counts.get("tp_linear_backward_weight", 0) == 0
    for p in g.nerf.mlp_feat.parameters():
        assert p.grad is None
    for p in list(g.nerf.mlp_rgb.parameters()) + list(g.nerf.mlp_trans.parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0
    assert g.latent_vars_trans.weight.grad is not None and g.latent_vars_light.weight.grad is not None


def test_fp32_mode_also_leaves_the_trunk_without_gradients():
    opt = adapt_gan_opt(H=32, W=32, sample_intvs=16, device=DEV)
    opt.b200 = AttrDict(mlp="fp32")
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=2).to(DEV)
    for p in g.nerf.parameters():
        p.requires_grad_(True)
    gen = torch.Generator().manual_seed(1)
    center = (torch.randn(2, 9, 3, generator=gen) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(DEV)
    ray = (torch.randn(2, 9, 3, generator=gen) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
    depth = ((torch.rand(2, 9, 16, 1, generator=gen) + torch.arange(16)[None, None, :, None]) / 16 * 1.2 + 0.2).to(DEV)
    out = g.nerf.forward_samples(opt, center, ray, depth, g.latent_vars_trans.weight, g.latent_vars_light.weight, mode="train")
    (out[0].mean() + out[1].mean() + out[2].mean()).backward()      # the static density has no gradient path either
    assert all(p.grad is None for p in g.nerf.mlp_feat.parameters())
    assert all(p.grad is not None for p in g.nerf.mlp_rgb.parameters())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_val_mode_broadcasts_latent_row_zero_to_every_view(precision):
    """mode='val' uses weight[0][None] (:592-593) for ALL views of the batch (the reference expands the
    single row); views 1.. must not read the latents of training images 1..  Row 1 is made very different from row 0, and
    the table has only two rows (reading "row 2" would be out of bounds), so a per-image read shows up as a large error.
    (Batched and single-view calls invert their pose matrices in differently shaped torch calls: a last-bit difference in a
    ray is amplified by the positional encoding, hence a tolerance instead of bit equality.)"""
    H, W, N = 16, 32, 32
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.nerf.sample_stratified = False
    opt.b200 = AttrDict(mlp=precision)
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=2).to(DEV)
    with torch.no_grad():
        g.latent_vars_trans.weight[1] = -4.0 * g.latent_vars_trans.weight[0] + 3.0
        g.latent_vars_light.weight[1] = -4.0 * g.latent_vars_light.weight[0] + 3.0
    B = 3
    pose = synth.poses([0, 1, 2]).to(DEV)
    intr = synth.intrinsics(B).clone()
    intr[:, :2] *= 0.05
    intr = intr.to(DEV)
    lo, hi = [t.to(DEV) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
    idx = torch.arange(H * W, device=DEV)[None].expand(B, -1)
    dr = (zn[:, :, None], zf[:, :, None])
    tol = 2e-3 if precision == "fp32" else 2e-2      # relative to the quantity's range (uncert reaches ~2 with these latents)
    with torch.no_grad():
        full = g.render(opt, pose, intr=intr, ray_idx=idx, depth_range=dr, mode="val")
        for b in range(B):
            one = g.render(opt, pose[b:b + 1], intr=intr[b:b + 1], ray_idx=idx[:1], depth_range=(dr[0][b:b + 1], dr[1][b:b + 1]),
                           mode="val")
            for k in ("rgb", "uncert", "rgb_transient"):
                scale = max(1.0, float(one[k].abs().max()))
                assert (full[k][b] - one[k][0]).abs().max() <= tol * scale, (b, k, float((full[k][b] - one[k][0]).abs().max()))
        # the same render with row 1's latents differs by far more than the tolerance: the check above is discriminating
        g.latent_vars_trans.weight[0], g.latent_vars_light.weight[0] = g.latent_vars_trans.weight[1].clone(), g.latent_vars_light.weight[1].clone()
        other = g.render(opt, pose[1:2], intr=intr[1:2], ray_idx=idx[:1], depth_range=(dr[0][1:2], dr[1][1:2]), mode="val")
        assert (other["rgb"][0] - full["rgb"][1]).abs().max() > 10 * tol


def test_every_loss_term_is_differentiable_and_reference_summarize_loss_works():
    """gradients arriving through render / uncert / trans_reg individually (what the reference's
    summarize_loss produces: all = sum 10^w * term built with torch ops) equal the oracle's autograd."""
    B, P, N, H, W = 4, 8, 16, 64, 64
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.batch_size, opt.patch_size = B, P
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=B).to(DEV)
    gen = torch.Generator().manual_seed(5)
    image = torch.rand(B, 3, H, W, generator=gen)
    mask = (torch.rand(B, H, W, generator=gen) > 0.5).float()
    coords, _ = synth.patch_coords(B, P, seed=4)
    rgb = torch.rand(B, P * P, 3, generator=gen)
    unc = torch.rand(B, P * P, 1, generator=gen) + 0.05
    dens = torch.rand(B, P * P, N, 2, generator=gen) * 4

    def ours(combine):
        r, u, d = [t.to(DEV).requires_grad_(True) for t in (rgb, unc, dens)]
        var = AttrDict(idx=torch.arange(B, device=DEV), image=image.to(DEV), obj_mask=mask.to(DEV), ray_idx=coords.to(DEV),
                       rgb=r, uncert=u, density=d)
        loss = g.compute_loss(opt, var, mode="train")
        assert "all" not in loss and set(loss) == {"render", "uncert", "trans_reg"}
        combine(loss, var).backward()
        return r.grad.cpu(), u.grad.cpu(), d.grad.cpu()

    def oracle(wr, wu, wt):
        r, u, d = [t.clone().requires_grad_(True) for t in (rgb, unc, dens)]
        ref = O.patch_losses(image, mask, coords, r, u, d, 0.0, 0.0, 0.0)      # unit weights: the raw terms
        (wr * ref["render"] + wu * ref["uncert"] + wt * ref["trans_reg"]).backward()
        return r.grad, u.grad, d.grad

    def reference_summarize(loss, var):          # model/base.py:145-157, verbatim arithmetic
        loss_all = 0.
        for key in loss:
            if opt.loss_weight[key] is not None:
                loss_all += 10 ** float(opt.loss_weight[key]) * loss[key]
        return loss_all

    cases = [(reference_summarize, (1.0, 1.0, 0.01)),
             (lambda l, v: base.summarize_loss(opt, v, l)["all"], (1.0, 1.0, 0.01)),
             (lambda l, v: 3.0 * l["render"], (3.0, 0.0, 0.0)),
             (lambda l, v: l["uncert"] - 2.0 * l["trans_reg"], (0.0, 1.0, -2.0)),
             (lambda l, v: base.summarize_loss(opt, v, l)["all"] + 0.5 * l["render"], (1.5, 1.0, 0.01))]
    for combine, w in cases:
        got, want = ours(combine), oracle(*w)
        for a, b in zip(got, want):
            assert (a - b).abs().max() <= 1e-5 * max(b.abs().max().item(), 1e-12), w
