"""Graph.render / render_by_slices through the drop-in module vs the reference fixture (train mode) and the
oracle (val / eval modes, which the reference can only run on CUDA)."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import compute_box, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model.nerf_adapt_st_gan import Graph
from tests.conftest import layer_list

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _graph(opt, g=None, n=4):
    torch.manual_seed(0)
    gr = Graph(opt)
    gr.latent_vars_trans = torch.nn.Embedding(n, 16)
    gr.latent_vars_light = torch.nn.Embedding(n, 48)
    torch.manual_seed(3)
    torch.nn.init.normal_(gr.latent_vars_trans.weight)
    torch.nn.init.normal_(gr.latent_vars_light.weight)
    if g is not None:
        assert torch.equal(gr.latent_vars_trans.weight.detach(), g.emb_trans)
    return gr.to(DEV)


@pytest.fixture(autouse=True)
def _reference_bit_constants():
    """K^-1 / pose^-1 as the CPU reference computes them: the 2^9*pi positional encoding amplifies ulp-level ray
    differences to ~1e-3 in the outputs, so end-to-end parity is checked on bit-identical rays."""
    from texpose_b200 import camera
    camera.HOST_MATRICES = True
    yield
    camera.HOST_MATRICES = False


def test_render_train_matches_reference(golden):
    g = golden("render_train")
    opt = adapt_gan_opt(H=128, W=128, device=DEV)
    gr = _graph(opt, g)
    torch.manual_seed(21)
    expect = torch.rand(2, 64, 64, 1, device=DEV)
    torch.manual_seed(21)
    ret = gr.render(opt, g.pose.to(DEV), intr=g.intr.to(DEV), ray_idx=g.coords.to(DEV),
                    depth_range=(g.z_near.to(DEV)[:, :, None], g.z_far.to(DEV)[:, :, None]),
                    sample_idx=g.sample_idx.to(DEV), mode="train")
    assert set(ret.keys()) == {"rgb", "rgb_static", "rgb_transient", "opacity", "opacity_static", "opacity_transient",
                               "uncert", "depth", "alpha_static", "alpha_transient", "density"}
    # the fixture's jitter came from the CPU generator: re-render with it injected for the value comparison
    if not torch.equal(expect.cpu(), g.rand):
        center, ray = gr.ray_sampler.get_rays(opt, g.intr.to(DEV), g.coords.to(DEV), g.pose.to(DEV))
        zn, zf = gr.ray_sampler.get_bounds(opt, g.coords.to(DEV), g.z_near.to(DEV), g.z_far.to(DEV))
        from texpose_b200 import ops
        depth = ops.sample_depth(zn.reshape(2, 64), zf.reshape(2, 64), 64, rand=g.rand.to(DEV))
        lt = gr.latent_vars_trans.weight[g.sample_idx.to(DEV)]
        ll = gr.latent_vars_light.weight[g.sample_idx.to(DEV)]
        rgb_s, dens, unc = gr.nerf.forward_samples(opt, center.view(2, 64, 3), ray.view(2, 64, 3), depth, lt, ll, "train")
        comp = gr.nerf.composite(opt, ray.view(2, 64, 3), rgb_s, dens, depth, unc)
        ret = AttrDict(rgb=comp[0], rgb_static=comp[1], rgb_transient=comp[2], depth=comp[3], opacity=comp[4],
                       opacity_static=comp[5], opacity_transient=comp[6], uncert=comp[8], alpha_static=comp[9],
                       alpha_transient=comp[10], density=dens)
    for k, v in ret.items():
        assert v.shape == g["o_" + k].shape, k
        assert (v.cpu() - g["o_" + k]).abs().max() <= TOL, (k, (v.cpu() - g["o_" + k]).abs().max())


def test_render_by_slices_val_and_eval_vs_oracle():
    H, W, N = 24, 32, 32
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.nerf.sample_stratified = False          # RNG-free
    gr = _graph(opt)
    pose = synth.poses([0])
    intr = synth.intrinsics(1).clone()
    intr[:, :2] *= 0.05
    lo, hi = synth.padded_aabb()
    zn, zf = compute_box.box_range(pose.to(DEV), intr.to(DEV), lo.to(DEV), hi.to(DEV), H, W, *synth.BG_RANGE)
    c, r = O.get_center_and_ray(pose, intr, H, W)
    tn, tf, v = O.aabb_ray_intersection(lo, hi, c, r)
    ozn, ozf = O.box_bounds_to_range(tn, tf, v, *synth.BG_RANGE)
    assert torch.equal(zn.cpu(), ozn) and torch.equal(zf.cpu(), ozf) and v.sum() > 20
    mask = v.view(1, H, W).float()
    var = AttrDict(pose=pose.to(DEV), intr=intr.to(DEV), z_near=zn, z_far=zf, obj_mask=mask.to(DEV),
                   idx=torch.tensor([0], device=DEV))
    # --- val: every pixel, latent row 0, chunks concatenated
    opt.b200 = AttrDict(slice_rays=300)         # force several slices
    with torch.no_grad():
        out = gr.nerf_forward(opt, AttrDict(var), mode="val")
    idx = torch.arange(H * W)[None]
    ref = O.render_stl(*( [O.gather_rays(t, idx) for t in O.get_center_and_ray(pose, intr, H, W)] ), ozn, ozf, None, N,
                       gr.latent_vars_trans.weight[0][None].cpu(), gr.latent_vars_light.weight[0][None].cpu(),
                       [(w.cpu(), b.cpu()) for w, b in layer_list(gr.nerf.mlp_feat)],
                       [(w.cpu(), b.cpu()) for w, b in layer_list(gr.nerf.mlp_rgb)],
                       [(w.cpu(), b.cpu()) for w, b in layer_list(gr.nerf.mlp_trans)])
    for k, vref in ref.items():
        assert out[k].shape == vref.shape, k
        assert (out[k].cpu() - vref).abs().max() <= TOL, k
    # --- eval: mask prior, defaults elsewhere, zero transient latent, picked light latent
    var.pose_anchor = synth.poses([0, 1, 2, 3]).to(DEV)
    opt.render.N_candidate = 1
    with torch.no_grad():
        out = gr.nerf_forward(opt, AttrDict(var), mode="eval_noalign")
    obj = v[0].nonzero()[:, 0]
    refm = O.render_stl(*( [O.gather_rays(t, obj[None]) for t in O.get_center_and_ray(pose, intr, H, W)] ),
                        ozn[:, obj], ozf[:, obj], None, N, torch.zeros(1, 16),
                        gr.latent_vars_light.weight[0][None].cpu(),
                        [(w.cpu(), b.cpu()) for w, b in layer_list(gr.nerf.mlp_feat)],
                        [(w.cpu(), b.cpu()) for w, b in layer_list(gr.nerf.mlp_rgb)],
                        [(w.cpu(), b.cpu()) for w, b in layer_list(gr.nerf.mlp_trans)])
    bg = (~v[0]).nonzero()[:, 0]
    for k, vref in refm.items():
        assert out[k].shape[1] == H * W
        assert (out[k][:, obj.to(DEV)].cpu() - vref).abs().max() <= TOL, k
    assert (out["rgb"][:, bg.to(DEV)] == 0).all() and (out["uncert"][:, bg.to(DEV)] == 0.05).all()
    assert (out["density"][:, bg.to(DEV)] == 1).all() and (out["alpha_static"][:, bg.to(DEV)] == 1).all()
    assert (out["depth"][:, bg.to(DEV)] == 0).all()


def test_batched_eval_views_and_frame_epilogue():
    """the eval branch for B > 1 views equals B single-view calls (the reference is B = 1 only), the latent
    pick runs per view on the device, and the per-frame epilogue (maps + PSNR, no host sync) matches the oracle."""
    H, W, N = 24, 32, 32
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.nerf.sample_stratified = False
    gr = _graph(opt)
    seeds = [0, 2, 5]
    pose = synth.poses(seeds)
    intr = synth.intrinsics(len(seeds)).clone()
    intr[:, :2] *= 0.05
    lo, hi = synth.padded_aabb()
    zn, zf = compute_box.box_range(pose.to(DEV), intr.to(DEV), lo.to(DEV), hi.to(DEV), H, W, *synth.BG_RANGE)
    c, r = O.get_center_and_ray(pose, intr, H, W)
    _, _, v = O.aabb_ray_intersection(lo, hi, c, r)
    mask = v.view(len(seeds), H, W).float().to(DEV)
    anchors = synth.poses([5, 1, 0, 3]).to(DEV)
    opt.render.N_candidate = 1                       # nearest anchor: view 0 -> row 2, view 1 -> some row, view 2 -> row 0
    var = AttrDict(pose=pose.to(DEV), intr=intr.to(DEV), z_near=zn, z_far=zf, obj_mask=mask, pose_anchor=anchors,
                   idx=torch.arange(len(seeds), device=DEV))
    with torch.no_grad():
        out = gr.nerf_forward(opt, AttrDict(var), mode="eval_noalign")
        for b in range(len(seeds)):
            one = AttrDict(pose=var.pose[b:b + 1], intr=var.intr[b:b + 1], z_near=zn[b:b + 1], z_far=zf[b:b + 1],
                           obj_mask=mask[b:b + 1], pose_anchor=anchors, idx=var.idx[b:b + 1])
            ref = gr.nerf_forward(opt, one, mode="eval_noalign")
            for k in ("rgb", "rgb_static", "depth", "opacity", "uncert", "density", "alpha_static"):
                assert torch.equal(out[k][b:b + 1], ref[k]), (b, k)
        gen = torch.Generator().manual_seed(8)
        out.image = torch.rand(len(seeds), 3, H, W, generator=gen).to(DEV)
        ev = gr.evaluate_frame(opt, out)
    want = O.eval_frame(out.rgb_static.cpu(), out.depth.cpu(), out.image.cpu(), mask.cpu(), H, W, float(opt.nerf.depth.scale))
    for k in ("rgb_map", "depth_map", "image_masked"):
        assert torch.equal(ev[k].cpu(), want[k].contiguous()), k
    assert (ev.mse.cpu() - want["mse"]).abs().max() <= 1e-6 * want["mse"].max()
    assert (ev.psnr.cpu() - want["psnr"]).abs().max() <= 1e-4
