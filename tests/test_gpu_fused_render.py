"""Fused render launch (tp_render_fused_forward: rays + depths + view bias + MLP + compositing in one kernel,
model/nerf_adapt_st_gan.py:565-631) against the multi-kernel path it replaces and against the CPU oracle."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import _C, compute_box, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model.nerf_adapt_st_gan import Graph
from tests.conftest import layer_list

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient", "uncert",
        "alpha_static", "alpha_transient", "density")


def _setup(H, W, N, seeds=(0,), stratified=False, rng="torch", scale=None, **b200):
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.nerf.sample_stratified = stratified
    opt.b200 = AttrDict(mlp="bf16", rng=rng, **b200)
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=4).to(DEV).eval()
    B = len(seeds)
    pose, intr = synth.poses(list(seeds)).to(DEV), synth.intrinsics(B).to(DEV).clone()
    intr[:, :2] *= (W / 640 if scale is None else scale)
    lo, hi = [t.to(DEV) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
    return opt, g, pose, intr, zn, zf


def _unfused(opt):
    o = AttrDict(opt)
    o.b200 = AttrDict(opt.b200)
    o.b200.fused_render = False
    return o


@pytest.mark.parametrize("N", [32, 64, 128])
@pytest.mark.parametrize("stratified,rng", [(False, "torch"), (True, "torch"), (True, "philox")])
def test_fused_equals_multikernel_path(N, stratified, rng):
    """Same rays (bit-identical generation), same depths (same arithmetic / same Philox stream), same bf16 MLP; only the
    summation order of the compositing differs (row-per-sample scan instead of lane-owns-K-samples)."""
    H, W = 24, 40
    opt, g, pose, intr, zn, zf = _setup(H, W, N, seeds=(0, 3), stratified=stratified, rng=rng, scale=0.06)
    dr = (zn[:, :, None], zf[:, :, None])
    gen = torch.Generator().manual_seed(5)
    idx = torch.stack([torch.randperm(H * W, generator=gen)[:333] for _ in range(2)]).to(DEV)      # ragged: 333 rays per view
    with torch.no_grad():
        g.render(opt, pose, intr=intr, ray_idx=range(0, 4), depth_range=dr, mode="val")      # first call packs the weight image
    for ray_idx in (idx, range(80, 80 + 7 * W + 3)):
        with torch.no_grad():
            _C.launch_counts.clear()
            torch.manual_seed(11)
            a = g.render(opt, pose, intr=intr, ray_idx=ray_idx, depth_range=dr, mode="val")
            assert _C.launch_counts.get("tp_render_fused_forward", 0) == 1 and "tp_composite_stl_forward" not in _C.launch_counts
            assert "tp_tc_ray_bias" not in _C.launch_counts and "tp_sample_depth" not in _C.launch_counts
            assert sum(_C.launch_counts.values()) == 3, dict(_C.launch_counts)      # two image-bias rows + the fused launch
            torch.manual_seed(11)
            b = g.render(_unfused(opt), pose, intr=intr, ray_idx=ray_idx, depth_range=dr, mode="val")
        for k in KEYS:
            assert a[k].shape == b[k].shape, k
            tol = 2e-5 * max(1.0, float(b[k].abs().max()))
            assert (a[k] - b[k]).abs().max() <= tol, (N, stratified, rng, k, float((a[k] - b[k]).abs().max()))
        assert torch.equal(a["density"], b["density"])          # per-sample head outputs: the same kernel arithmetic
        assert torch.isfinite(a["rgb"]).all()


def test_fused_eval_mode_mask_prior_and_static_only():
    H, W, N = 24, 32, 64
    opt, g, pose, intr, zn, zf = _setup(H, W, N, seeds=(0,), scale=0.05)
    c, r = O.get_center_and_ray(pose.cpu(), intr.cpu(), H, W)
    lo, hi = synth.padded_aabb()
    _, _, v = O.aabb_ray_intersection(lo, hi, c, r)
    var = AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=v.view(1, H, W).float().to(DEV),
                   idx=torch.zeros(1, dtype=torch.long, device=DEV), pose_anchor=synth.poses([0, 1, 2, 3]).to(DEV))
    opt.render.N_candidate = 1
    with torch.no_grad():
        a = g.nerf_forward(opt, AttrDict(var), mode="eval_noalign")
        b = g.nerf_forward(_unfused(opt), AttrDict(var), mode="eval_noalign")
        opt.b200.static_only = True
        s = g.nerf_forward(opt, AttrDict(var), mode="eval")
    for k in KEYS:
        assert (a[k] - b[k]).abs().max() <= 2e-5 * max(1.0, float(b[k].abs().max())), k
    for k in ("rgb_static", "depth", "opacity_static"):        # what Model.evaluate_full reads (:341-362)
        assert (s[k] - a[k]).abs().max() <= 2e-5 * max(1.0, float(a[k].abs().max())), k
    obj = v[0].to(DEV)
    assert float(s["rgb_transient"][:, obj].abs().max()) == 0.0 and float(s["density"][:, obj][..., 1].abs().max()) == 0.0


def test_c2_config_rays_vs_oracle_all_outputs():
    """BASELINE configs[1] -- the 480x640 frame at 128 samples per ray, bf16 tensor-core MLP -- on
    the frame's AABB-bounded rays (>= 1024 of them, strided over the object so every image row block is hit), all eleven
    outputs of Graph.render against the CPU oracle on the same inputs.  Tolerance: north-star 1e-2 max-abs (bf16 path) on
    the rendered quantities; raw densities (unbounded softplus outputs) within 2 % of their range."""
    H, W, N = 480, 640, 128
    opt, g, pose, intr, zn, zf = _setup(H, W, N, seeds=(0,))
    c, r = O.get_center_and_ray(pose.cpu(), intr.cpu(), H, W)
    lo, hi = synth.padded_aabb()
    tn, tf, v = O.aabb_ray_intersection(lo, hi, c, r)
    ozn, ozf = O.box_bounds_to_range(tn, tf, v, *synth.BG_RANGE)
    assert torch.equal(zn.cpu(), ozn) and torch.equal(zf.cpu(), ozf)
    obj = v[0].nonzero()[:, 0]
    pick = obj[:: max(1, len(obj) // 1536)][:1536]
    assert len(pick) >= 1024
    with torch.no_grad():
        got = g.render(opt, pose, intr=intr, ray_idx=pick[None].to(DEV), depth_range=(zn[:, :, None], zf[:, :, None]), mode="val")
    L = lambda ml: [(w.cpu(), b.cpu()) for w, b in layer_list(ml)]
    ref = O.render_stl(O.gather_rays(c, pick[None]), O.gather_rays(r, pick[None]), ozn[:, pick], ozf[:, pick], None, N,
                       g.latent_vars_trans.weight[0][None].detach().cpu(), g.latent_vars_light.weight[0][None].detach().cpu(),
                       L(g.nerf.mlp_feat), L(g.nerf.mlp_rgb), L(g.nerf.mlp_trans))
    errs = {k: float((got[k].cpu() - ref[k]).abs().max()) for k in KEYS}
    print("C2 rays, fused bf16 render vs oracle:", {k: f"{e:.2e}" for k, e in errs.items()})
    for k in ("rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient",
              "alpha_static", "alpha_transient"):
        assert errs[k] <= 1e-2, (k, errs[k])
    assert errs["uncert"] <= 1.5e-2, errs["uncert"]      # unbounded quantity (reaches ~1.8): < 1 % of its range, trunk-bound
    assert errs["density"] <= 0.02 * float(ref["density"].abs().max()), errs["density"]


def test_full_frame_fused_properties_and_row_shards():
    """The whole C2 frame in one fused launch: deterministic, opacity == 1 (1e10 tail), and 8-way row-block shards (what 8
    GPUs each render in the one-frame strong-scaling mode) equal the whole frame bit for bit."""
    from texpose_b200 import parallel
    H, W, N = 480, 640, 128
    opt, g, pose, intr, zn, zf = _setup(H, W, N, seeds=(0,))
    dr = (zn[:, :, None], zf[:, :, None])
    var = AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=torch.ones(1, H, W, device=DEV),
                   idx=torch.zeros(1, dtype=torch.long, device=DEV))
    with torch.no_grad():
        g.render(opt, pose, intr=intr, ray_idx=range(0, 4), depth_range=dr, mode="val")      # first call packs the weight image
        _C.launch_counts.clear()
        full = g.nerf_forward(opt, AttrDict(var), mode="val")
        assert sum(_C.launch_counts.values()) == 3 and _C.launch_counts["tp_render_fused_forward"] == 1
        again = g.nerf_forward(opt, AttrDict(var), mode="val")
    for k in ("rgb", "depth", "uncert", "opacity", "alpha_static"):
        assert torch.equal(full[k], again[k]) and torch.isfinite(full[k]).all(), k
    assert full.rgb.shape == (1, H * W, 3) and full.density.shape == (1, H * W, N, 2) and full.alpha_static.shape == (1, H * W, N)
    assert (full.opacity - 1).abs().max() < 1e-4 and (full.opacity_static - 1).abs().max() < 1e-4
    assert (full.depth >= zn[:, :, None] - 1e-3).all() and (full.depth <= zf[:, :, None] + 1e-3).all()
    for rank in (0, 3, 7):
        b, e = parallel.shard_rays(H * W, rank, 8, align=W)
        with torch.no_grad():
            part = g.render(opt, pose, intr=intr, ray_idx=range(b, e), depth_range=dr, mode="val")
        for k in ("rgb", "depth", "uncert", "opacity", "rgb_static", "alpha_transient"):
            assert torch.equal(part[k], full[k][:, b:e]), (rank, k)


def test_fused_outputs_can_be_skipped_and_redirected():
    """want= trims the per-sample tensors (inference that only needs the image); out_ptrs= sends an output to a caller-owned
    address (the multi-GPU frame gather points it at a peer window)."""
    H, W, N = 16, 32, 128
    opt, g, pose, intr, zn, zf = _setup(H, W, N, seeds=(1,), scale=0.05)
    dr = (zn[:, :, None], zf[:, :, None])
    with torch.no_grad():
        full = g.render(opt, pose, intr=intr, ray_idx=range(0, H * W), depth_range=dr, mode="val")
        frame = torch.full((H * W, 3), -1.0, device=DEV)
        b, e = 3 * W, 9 * W
        part = g._render_fused(opt, pose, intr, range(b, e), dr, None, "val", want=("rgb", "depth"),
                               out_ptrs={"rgb": frame.data_ptr() + b * 3 * 4})
    assert set(part) == {"depth"} and torch.equal(part["depth"], full["depth"][:, b:e])
    assert torch.equal(frame[b:e], full["rgb"][0, b:e]) and float(frame[:b].max()) == -1.0 and float(frame[e:].max()) == -1.0
