"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference checkout does not exist on
the GPU box):   python -m oracle.make_golden
Every fixture stores the inputs next to the reference's outputs, so the tests never need the
reference at run time.  Network weights are NOT stored (3.6 MB): they are re-created from
torch.manual_seed(seed) by constructing the layers in the reference's order, and each fixture
carries checksums that pin that re-creation.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_import
from texpose_b200 import synth
from texpose_b200.config import AttrDict

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def weight_checksums(module):
    out = {}
    for name, p in module.named_parameters():
        if p.dim() == 0:
            continue
        out["ck_sum/" + name] = p.detach().double().sum().item()
        out["ck_abs/" + name] = p.detach().double().abs().sum().item()
        out["ck_head/" + name] = p.detach().flatten()[:4].clone()
    return out


def case_rays(ns, opt):
    """camera.get_center_and_ray + aabb_ray_intersection on a strided pixel subset of a 480x640 frame."""
    pose = synth.poses([0, 1])
    intr = synth.intrinsics(2)
    center, ray = ns.camera.get_center_and_ray(opt, pose, intr=intr)
    lo, hi = synth.padded_aabb()
    t_near, t_far, valid = ns.camera.aabb_ray_intersection(lo, hi, center, ray)
    # keep every 53rd pixel + the whole silhouette band of view 0 rows 200..203
    idx = torch.cat([torch.arange(0, 480 * 640, 53), torch.arange(200 * 640, 204 * 640)]).unique()
    # axis-parallel rays (ray component exactly 0 -> inf / NaN handling of the slab test)
    ro = torch.tensor([[[0.0, 0.0, -5.0], [0.5, 0.0, -5.0], [2.0, 0.0, -5.0], [0.0, 0.0, 5.0],
                        [0.87, 0.1, -5.0], [0.0, 0.0, 0.0]]])
    rd = torch.tensor([[[0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0],
                        [0.0, 0.0, 1.0], [0.0, 1.0, 0.0]]])
    sn, sf, sv = ns.camera.aabb_ray_intersection(lo, hi, ro, rd)
    return dict(pose=pose, intr=intr, H=480, W=640, idx=idx, center=center[:, idx], ray=ray[:, idx],
                aabb_min=lo, aabb_max=hi, t_near=t_near[:, idx], t_far=t_far[:, idx], valid=valid[:, idx],
                n_valid=valid.sum(dim=1), special_o=ro, special_d=rd, special_near=sn, special_far=sf,
                special_valid=sv)


def case_patch(ns, opt128):
    """RaySampler.get_rays / get_bounds on 128x128 crops."""
    B, P = 3, 8
    coords, _ = synth.patch_coords(B, P, seed=2)
    pose = synth.poses([3, 4, 5])
    intr = crop_intrinsics(B)
    c_full, r_full = ns.camera.get_center_and_ray(opt128, pose, intr=intr)
    lo, hi = synth.padded_aabb()
    tn, tf, valid = ns.camera.aabb_ray_intersection(lo, hi, c_full, r_full)
    z_near = torch.where(valid, tn, torch.full_like(tn, synth.BG_RANGE[0]))
    z_far = torch.where(valid, tf, torch.full_like(tf, synth.BG_RANGE[1]))
    center, ray = ns.ray_sampler.RaySampler.get_rays(opt128, intr, coords, pose)
    zn, zf = ns.ray_sampler.RaySampler.get_bounds(opt128, coords, z_near, z_far)
    return dict(coords=coords, pose=pose, intr=intr, H=128, W=128, z_near=z_near, z_far=z_far,
                center=center, ray=ray, zn=zn, zf=zf)


def crop_intrinsics(B):
    """128x128 crop whose principal point keeps the object (t = OBJ_T) near the image centre."""
    K = torch.tensor([[572.4114, 0, 64 - 572.4114 * 0.3 / 8], [0, 573.57043, 64 + 573.57043 * 0.2 / 8], [0, 0, 1]])
    return K.repeat(B, 1, 1).float()


def case_sample_depth(ns, opt):
    out = {}
    torch.manual_seed(5)
    zn = torch.rand(2, 37) * 3 + 5
    zf = zn + torch.rand(2, 37) * 4 + 0.1
    for n in (64, 128, 48):
        o = AttrDict(opt)
        o.nerf = AttrDict(opt.nerf)
        o.nerf.sample_intvs = n
        torch.manual_seed(11)
        rand = torch.rand(2, 37, n, 1)
        torch.manual_seed(11)
        d = ns.graph_mod.Graph.sample_depth(o, 2, (zn, zf), num_rays=37)
        out[f"rand{n}"] = rand
        out[f"depth{n}"] = d
    o = AttrDict(opt)
    o.nerf = AttrDict(opt.nerf)
    o.nerf.sample_stratified = False
    out["depth64_mid"] = ns.graph_mod.Graph.sample_depth(o, 2, (zn, zf), num_rays=37)
    out.update(z_near=zn, z_far=zf)
    return out


def valid_rays(ns, opt, seed_pose, count):
    pose = synth.poses([seed_pose])
    intr = synth.intrinsics(1)
    center, ray = ns.camera.get_center_and_ray(opt, pose, intr=intr)
    lo, hi = synth.padded_aabb()
    tn, tf, valid = ns.camera.aabb_ray_intersection(lo, hi, center, ray)
    idx = valid[0].nonzero()[:, 0]
    idx = idx[torch.linspace(0, len(idx) - 1, count).long()]
    return center[:, idx], ray[:, idx], tn[:, idx], tf[:, idx], idx


def case_nerf_stl(ns, opt):
    """forward_samples + composite + backward of the three loss seeds (static/transient/light NeRF)."""
    R, N = 48, 64
    torch.manual_seed(0)
    nerf = ns.nerf_stl.NeRF(opt)
    center, ray, tn, tf, idx = valid_rays(ns, opt, 0, R)
    g = torch.Generator().manual_seed(7)
    rand = torch.rand(1, R, N, 1, generator=g)
    depth = (rand + torch.arange(N)[None, None, :, None].float()) / N * (tf - tn)[:, :, None, None] + tn[:, :, None, None]
    lt, ll = synth.latents(1, seed=1)
    lt = lt.clone().requires_grad_(True)
    ll = ll.clone().requires_grad_(True)
    rgb_s, dens, unc = nerf.forward_samples(opt, center, ray, depth, latent_variable_trans=lt,
                                            latent_variable_light=ll, mode="train")
    comp = nerf.composite(opt, ray, rgb_s, dens, depth, unc)
    names = ["rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient",
             "prob", "uncert", "alpha_static", "alpha_transient"]
    out = {"o_" + k: v for k, v in zip(names, comp)}
    # the three ray-wise loss terms of compute_loss (model/nerf_adapt_st_gan.py:747-763)
    image = torch.rand(1, R, 3, generator=g)
    mask = (torch.rand(1, R, 1, generator=g) > 0.3).float()
    rgb, uncert = comp[0], comp[8]
    loss = (mask * ((image - rgb) ** 2 / uncert ** 2)).sum() / (mask.sum() + 1e-5) \
        + (5 + torch.log(uncert ** 2).mean() / 2) + 0.01 * dens[..., -1].mean()
    loss.backward()
    out.update(center=center, ray=ray, depth=depth, latent_trans=lt, latent_light=ll, image=image, mask=mask,
               rgb_samples=rgb_s, density_samples=dens, uncert_samples=unc, loss=loss,
               g_latent_trans=lt.grad, g_latent_light=ll.grad)
    for name, p in nerf.named_parameters():
        if p.grad is not None:
            out["g/" + name] = p.grad if p.grad.numel() <= 2048 else p.grad[:, ::8][::8].contiguous()
            out["gsum/" + name] = p.grad.double().sum().item()
            out["gabs/" + name] = p.grad.double().abs().sum().item()
    out.update(weight_checksums(nerf))
    out["posenc_x"] = nerf.positional_encoding(opt, center[0, :5], L=10)
    return out


def case_render_train(ns, opt128):
    """Graph.render(mode='train') end to end on CPU (patch rays -> bounds -> depths -> MLP -> composite)."""
    B, P = 2, 8
    g = ref_import.build_graph(ns, opt128, n_images=4, seed=0)
    torch.manual_seed(3)
    torch.nn.init.normal_(g.latent_vars_trans.weight)
    torch.nn.init.normal_(g.latent_vars_light.weight)
    coords, _ = synth.patch_coords(B, P, seed=9)
    pose = synth.poses([6, 7])
    intr = crop_intrinsics(B)
    c_full, r_full = ns.camera.get_center_and_ray(opt128, pose, intr=intr)
    lo, hi = synth.padded_aabb()
    tn, tf, valid = ns.camera.aabb_ray_intersection(lo, hi, c_full, r_full)
    z_near = torch.where(valid, tn, torch.full_like(tn, synth.BG_RANGE[0]))
    z_far = torch.where(valid, tf, torch.full_like(tf, synth.BG_RANGE[1]))
    sample_idx = torch.tensor([2, 1])
    torch.manual_seed(21)
    rand = torch.rand(B, P * P, opt128.nerf.sample_intvs, 1)
    torch.manual_seed(21)
    ret = g.render(opt128, pose, intr=intr, ray_idx=coords, depth_range=(z_near[:, :, None], z_far[:, :, None]),
                   sample_idx=sample_idx, mode="train")
    out = {"o_" + k: v for k, v in ret.items()}
    out.update(coords=coords, pose=pose, intr=intr, z_near=z_near, z_far=z_far, sample_idx=sample_idx, rand=rand,
               emb_trans=g.latent_vars_trans.weight, emb_light=g.latent_vars_light.weight, H=128, W=128)
    return out


def case_loss(ns, opt128):
    """Graph.compute_loss(train_step='nerf') + Model.summarize_loss of the reference (model/nerf_adapt_st_gan.py:712-763,
    model/base.py:145-157) with the yaml's ray-wise terms (render, uncert, trans_reg); the VGG / GAN / Lab terms are switched
    off (loss_weight None) -- they need pretrained nets and are outside the path.  Also FlexPatchSampler (seeded)."""
    import copy
    from easydict import EasyDict as edict
    opt = copy.deepcopy(opt128)
    opt.loss_weight.feat = None
    opt.loss_weight.gan_nerf = None
    opt.loss_weight.lab = None
    opt.gan = None
    B, P, N = 3, 8, 16
    g = ref_import.build_graph(ns, opt, n_images=4, seed=0)
    gen = torch.Generator().manual_seed(31)
    image = torch.rand(B, 3, 128, 128, generator=gen)
    obj_mask = (torch.rand(B, 128, 128, generator=gen) > 0.4).float() * 255.0      # masks arrive as 0 / 255 maps
    coords, _ = synth.patch_coords(B, P, seed=5)
    coords[1] = coords[1] * 1.02                                                     # a few samples land outside [-1, 1]
    rgb = torch.rand(B, P * P, 3, generator=gen).requires_grad_(True)
    uncert = (torch.rand(B, P * P, 1, generator=gen) * 0.8 + 0.05).requires_grad_(True)
    density = (torch.rand(B, P * P, N, 2, generator=gen) * 3).requires_grad_(True)
    var = edict(idx=torch.arange(B), rgb=rgb, uncert=uncert, density=density, image=image, obj_mask=obj_mask,
                ray_idx=coords, opacity=torch.rand(B, P * P, 1, generator=gen))
    loss = g.compute_loss(opt, var, mode="train", train_step="nerf")
    base = __import__("importlib").import_module("model.base")
    loss = base.Model.summarize_loss(None, opt, var, loss)
    loss.all.backward()
    torch.manual_seed(77)
    sampler = ns.patch_sampler.FlexPatchSampler(random_shift=True, random_scale=True, min_scale=0.25, max_scale=1.0)
    fc, fs = sampler(nbatch=5, patch_size=6, device="cpu")
    return dict(image=image, obj_mask=obj_mask, coords=coords, rgb=rgb, uncert=uncert, density=density, H=128, W=128,
                image_sample=var.image_sample, mask_sample=var.mask_sample, l_render=loss.render, l_uncert=loss.uncert,
                l_trans_reg=loss.trans_reg, l_all=loss["all"], g_rgb=rgb.grad, g_uncert=uncert.grad, g_density=density.grad,
                w_render=opt.loss_weight.render, w_uncert=opt.loss_weight.uncert, w_trans_reg=opt.loss_weight.trans_reg,
                flex_coords=fc, flex_scales=fs, flex_seed=77)


def case_plain(ns):
    """layers/nerf.py NeRF (trunk trainable) with the nerf_lm_env.yaml dims: forward, composite, grads."""
    opt = ref_import.load_yaml_opt("nerf_lm_env", H=480, W=640)
    R, N = 24, 64
    torch.manual_seed(0)
    nerf = ns.nerf_plain.NeRF(opt)
    optf = ref_import.load_yaml_opt("nerf_lm_adapt_gan")
    center, ray, tn, tf, idx = valid_rays(ns, optf, 2, R)
    g = torch.Generator().manual_seed(8)
    rand = torch.rand(1, R, N, 1, generator=g)
    depth = (rand + torch.arange(N)[None, None, :, None].float()) / N * (tf - tn)[:, :, None, None] + tn[:, :, None, None]
    rgb_s, dens = nerf.forward_samples(opt, center, ray, depth, mode="train")
    rgb, d, op, prob = nerf.composite(opt, ray, rgb_s, dens, depth)
    image = torch.rand(1, R, 3, generator=g)
    loss = ((rgb - image) ** 2).mean() + 0.1 * ((d - 8.0) ** 2).mean() + 0.05 * op.mean()
    loss.backward()
    out = dict(center=center, ray=ray, depth=depth, rgb_samples=rgb_s, density_samples=dens, o_rgb=rgb, o_depth=d,
               o_opacity=op, o_prob=prob, image=image, loss=loss)
    for name, p in nerf.named_parameters():
        if p.grad is not None:
            out["g/" + name] = p.grad if p.grad.numel() <= 2048 else p.grad[:, ::8][::8].contiguous()
            out["gsum/" + name] = p.grad.double().sum().item()
            out["gabs/" + name] = p.grad.double().abs().sum().item()
    out.update(weight_checksums(nerf))
    return out


def case_pretrain(ns):
    """model/nerf_pretrain.py Graph end to end on CPU: forward(mode='train') (get_ray_idx -> render) + compute_loss + autograd,
    and render_by_slices(mode='val') over a whole 96 x 128 frame.  The reference hard-codes `.cuda()` in ray_batch_sample (:462);
    for this CPU run Tensor.cuda is the identity while the case runs (the reference itself is not modified)."""
    import importlib
    mod = importlib.import_module("model.nerf_pretrain")
    H, W, N, B = 96, 128, 32, 2
    opt = ref_import.load_yaml_opt("nerf_lm_env", H=H, W=W)
    opt.nerf.sample_intvs, opt.nerf.rand_rays = N, 512
    opt.loss_weight.update(render=0, mask=-1, depth=-1)
    opt.data.erode_mask_loss = False
    torch.manual_seed(0)
    g = mod.Graph(opt)
    pose = synth.poses([0, 1])
    intr = synth.intrinsics(B).clone()
    intr[:, :2] *= 0.2
    c_full, r_full = ns.camera.get_center_and_ray(opt, pose, intr=intr)
    lo, hi = synth.padded_aabb()
    tn, tf, valid = ns.camera.aabb_ray_intersection(lo, hi, c_full, r_full)
    z_near = torch.where(valid, tn, torch.full_like(tn, 7.0))
    z_far = torch.where(valid, tf, torch.full_like(tf, 9.0))
    gen = torch.Generator().manual_seed(3)
    image = torch.rand(B, 3, H, W, generator=gen)
    depth_gt = 7.5 + torch.rand(B, H, W, generator=gen)
    obj_mask = valid.view(B, H, W).float()
    from texpose_b200.config import AttrDict
    var = AttrDict(idx=torch.arange(B), pose=pose, pose_init=pose, intr=intr, z_near=z_near, z_far=z_far, image=image,
                   obj_mask=obj_mask, depth_gt=depth_gt)
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        torch.manual_seed(21)
        ray_idx = torch.randperm(H * W)[:opt.nerf.rand_rays // B].repeat(B, 1)
        rand = torch.rand(B, opt.nerf.rand_rays // B, N, 1)
        torch.manual_seed(21)
        var = g.forward(opt, var, mode="train")
        assert torch.equal(var.ray_idx, ray_idx)
        loss = g.compute_loss(opt, var, mode="train")
        total = sum(10 ** float(opt.loss_weight[k]) * loss[k] for k in loss)
        total.backward()
        out = dict(pose=pose, intr=intr, z_near=z_near, z_far=z_far, image=image, depth_gt=depth_gt, obj_mask=obj_mask, H=H, W=W,
                   N=N, ray_idx=ray_idx, rand=rand, o_rgb=var.rgb, o_depth=var.depth, o_opacity=var.opacity, l_total=total)
        out.update({"l_" + k: v for k, v in loss.items()})
        for name, p in g.nerf.named_parameters():
            out["g/" + name] = p.grad if p.grad.numel() <= 2048 else p.grad[:, ::8][::8].contiguous()
            out["gsum/" + name] = p.grad.double().sum().item()
            out["gabs/" + name] = p.grad.double().abs().sum().item()
        opt.nerf.sample_stratified = False
        with torch.no_grad():
            val = g.render_by_slices(opt, pose[:1], intr=intr[:1], depth_range=(z_near[:1, :, None], z_far[:1, :, None]),
                                     object_mask=obj_mask[:1], mode="val")
        out.update(v_rgb=val.rgb, v_depth=val.depth, v_opacity=val.opacity)
    finally:
        torch.Tensor.cuda = cuda
    out.update(weight_checksums(g.nerf))
    return out


def case_normals():
    box, surfel = ref_import.load_surfel()
    H, W = 120, 160
    pose = synth.poses([0, 5])
    intr = synth.intrinsics(2).clone()
    intr[:, :2] *= 0.25
    depth = synth.ellipsoid_depth(pose, intr, H, W, radii=[1.5, 1.1, 1.3])
    n = surfel.normal_from_depth(pose, depth, intr, H, W)
    zn = torch.where(depth * 0.8 > 0, depth * 0.8, torch.full_like(depth, synth.BG_RANGE[0]))   # data/lm.py:352-356
    zf = torch.where(depth * 1.2 > 0, depth * 1.2, torch.full_like(depth, synth.BG_RANGE[1]))
    return dict(pose=pose, intr=intr, depth=depth, normal=n, H=H, W=W, guided_near=zn, guided_far=zf)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(True)
    ns = ref_import.load()
    opt = ref_import.load_yaml_opt("nerf_lm_adapt_gan", H=480, W=640)
    opt128 = ref_import.load_yaml_opt("nerf_lm_adapt_gan", H=128, W=128)
    cases = dict(
        rays=lambda: case_rays(ns, opt),
        patch=lambda: case_patch(ns, opt128),
        sample_depth=lambda: case_sample_depth(ns, opt),
        nerf_stl=lambda: case_nerf_stl(ns, opt),
        render_train=lambda: case_render_train(ns, opt128),
        plain=lambda: case_plain(ns),
        normals=case_normals,
        loss=lambda: case_loss(ns, opt128),
        pretrain=lambda: case_pretrain(ns),
    )
    for name, fn in cases.items():
        d = _np(fn())
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name:14s} {os.path.getsize(path) / 1024:8.1f} KiB  {len(d)} arrays")


if __name__ == "__main__":
    main()
