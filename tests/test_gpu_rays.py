"""K1 parity on the GPU: rays, AABB, sample placement, patch rays/bounds, gather, box range, normals.
Bit-exact where the contract says so (a3-a8), through the C-ABI."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import camera, compute_box, compute_surfelinfo, ops, synth
from texpose_b200.config import adapt_gan_opt
from texpose_b200.model.nerf_adapt_st_gan import Graph
from texpose_b200.tools.ray_sampler import RaySampler

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_full_frame_rays_bit_exact_vs_golden(golden):
    g = golden("rays")
    opt = adapt_gan_opt(device=DEV)
    center, ray = camera.get_center_and_ray(opt, g.pose.to(DEV), intr=g.intr.to(DEV))
    assert center.shape == (2, 480 * 640, 3)
    # K^-1 and pose^-1 come from torch on the GPU here; compare against the reference (CPU) fixture
    assert (center[:, g.idx.to(DEV)].cpu() - g.center).abs().max() <= 2e-6
    assert (ray[:, g.idx.to(DEV)].cpu() - g.ray).abs().max() <= 2e-6
    # with the host constants computed exactly as the reference did (CPU), the rays are bit-identical
    kinv, pinv = camera.view_matrices(g.pose, g.intr)
    c2, r2 = ops.raygen(kinv.to(DEV), pinv.to(DEV), 480, 640, 0.5, None)
    assert torch.equal(c2[:, g.idx.to(DEV)].cpu(), g.center)
    assert torch.equal(r2[:, g.idx.to(DEV)].cpu(), g.ray)


def test_ray_idx_gather_fusion(golden):
    g = golden("rays")
    kinv, pinv = camera.view_matrices(g.pose, g.intr)
    idx = g.idx.to(DEV)[None].expand(2, -1).contiguous()
    c, r = ops.raygen(kinv.to(DEV), pinv.to(DEV), 480, 640, 0.5, idx)
    assert torch.equal(c.cpu(), g.center) and torch.equal(r.cpu(), g.ray)


def test_aabb_bit_exact(golden):
    g = golden("rays")
    tn, tf, valid = camera.aabb_ray_intersection(g.aabb_min.to(DEV), g.aabb_max.to(DEV), g.center.to(DEV), g.ray.to(DEV))
    assert torch.equal(tn.cpu(), g.t_near) and torch.equal(tf.cpu(), g.t_far) and torch.equal(valid.cpu(), g.valid)
    # axis-parallel rays: inf / NaN propagation identical to torch.minimum/maximum/max/min
    sn, sf, sv = camera.aabb_ray_intersection(g.aabb_min.to(DEV), g.aabb_max.to(DEV), g.special_o.to(DEV),
                                              g.special_d.to(DEV))
    assert torch.equal(sv.cpu(), g.special_valid)
    assert torch.equal(torch.nan_to_num(sn.cpu(), nan=-7.0), torch.nan_to_num(g.special_near, nan=-7.0))
    assert torch.equal(torch.nan_to_num(sf.cpu(), nan=-7.0), torch.nan_to_num(g.special_far, nan=-7.0))


def test_full_frame_valid_count_and_box_range(golden):
    g = golden("rays")
    kinv, pinv = camera.view_matrices(g.pose, g.intr)
    zn, zf, valid = ops.box_range(kinv.to(DEV), pinv.to(DEV), 480, 640, g.aabb_min.to(DEV), g.aabb_max.to(DEV),
                                  *synth.BG_RANGE, want_valid=True)
    assert torch.equal(valid.sum(dim=1).cpu(), g.n_valid)
    c, r = O.get_center_and_ray(g.pose, g.intr, 480, 640)
    tn, tf, v = O.aabb_ray_intersection(g.aabb_min, g.aabb_max, c, r)
    ozn, ozf = O.box_bounds_to_range(tn, tf, v, *synth.BG_RANGE)
    assert torch.equal(zn.cpu(), ozn) and torch.equal(zf.cpu(), ozf)
    bounds, _ = compute_box.box_bounds(g.pose.to(DEV), g.intr.to(DEV), g.aabb_min.to(DEV), g.aabb_max.to(DEV))
    assert bounds.shape == (2, 2, 480, 640)
    assert (bounds[:, 0].flatten(1).cpu() - torch.where(v, tn, torch.zeros_like(tn))).abs().max() <= 2e-5


def test_sample_depth_bit_exact(golden):
    g = golden("sample_depth")
    for n in (64, 128, 48):
        d = ops.sample_depth(g.z_near.to(DEV), g.z_far.to(DEV), n, rand=g[f"rand{n}"].to(DEV))
        assert torch.equal(d.cpu(), g[f"depth{n}"]), n
    d = ops.sample_depth(g.z_near.to(DEV), g.z_far.to(DEV), 64, stratified=False)
    assert torch.equal(d.cpu(), g.depth64_mid)


def test_sample_depth_philox_statistics():
    zn = torch.full((4, 1000), 2.0, device=DEV)
    zf = torch.full((4, 1000), 6.0, device=DEV)
    d = ops.sample_depth(zn, zf, 64, seed=1234)[..., 0]
    k = torch.arange(64, device=DEV, dtype=torch.float32)
    lo, hi = k / 64 * 4 + 2, (k + 1) / 64 * 4 + 2
    assert ((d >= lo) & (d <= hi)).all()                                        # sample i stays inside stratum i
    u = (d - 2.0) / 4.0 * 64 - k
    assert abs(u.mean().item() - 0.5) < 5e-3 and abs(u.var().item() - 1 / 12) < 5e-3
    d2 = ops.sample_depth(zn, zf, 64, seed=1234)[..., 0]
    assert torch.equal(d, d2)
    assert not torch.equal(d, ops.sample_depth(zn, zf, 64, seed=99)[..., 0])


def test_patch_rays_and_bounds(golden):
    g = golden("patch")
    opt = adapt_gan_opt(H=128, W=128, device=DEV)
    kinv, pinv = camera.view_matrices(g.pose, g.intr)
    c, r = ops.patch_rays(kinv.to(DEV), pinv.to(DEV), g.coords.to(DEV), 128, 128)
    assert torch.equal(c.cpu(), g.center) and torch.equal(r.cpu(), g.ray)      # bit-exact with reference-bit constants
    c2, r2 = RaySampler.get_rays(opt, g.intr.to(DEV), g.coords.to(DEV), g.pose.to(DEV))
    assert (r2.cpu() - g.ray).abs().max() <= 2e-6                               # K^-1 / pose^-1 from the GPU solver
    camera.HOST_MATRICES = True
    try:
        c3, r3 = RaySampler.get_rays(opt, g.intr.to(DEV), g.coords.to(DEV), g.pose.to(DEV))
    finally:
        camera.HOST_MATRICES = False
    assert torch.equal(r3.cpu(), g.ray)
    zn, zf = RaySampler.get_bounds(opt, g.coords.to(DEV), g.z_near.to(DEV), g.z_far.to(DEV))
    assert torch.equal(zn.cpu(), g.zn) and torch.equal(zf.cpu(), g.zf)          # torch CPU grid_sample arithmetic
    img = torch.rand(3, 5, 128, 128)
    ref = torch.nn.functional.grid_sample(img, g.coords, mode="bilinear", align_corners=True)
    got = RaySampler.get_image(opt, g.coords.to(DEV), img.to(DEV))
    assert (got.cpu() - ref).abs().max() <= 2e-6


def test_ray_batch_sample_bit_exact():
    x = torch.randn(3, 500, 4)
    idx = torch.randint(0, 500, (3, 77))
    assert torch.equal(Graph.ray_batch_sample(x.to(DEV), idx.to(DEV)).cpu(), O.gather_rays(x, idx))
    empty = Graph.ray_batch_sample(x.to(DEV), idx[:, :0].to(DEV))
    assert empty.shape == (3, 0, 4)


def test_points_from_depth_bit_exact(golden):
    g = golden("nerf_stl")
    opt = adapt_gan_opt(device=DEV)
    p = camera.get_3D_points_from_depth(opt, g.center.to(DEV), g.ray.to(DEV), g.depth.to(DEV), multi_samples=True)
    assert torch.equal(p.cpu(), O.points_from_depth(g.center, g.ray, g.depth))


def test_normals_and_guided_range(golden):
    g = golden("normals")
    n = compute_surfelinfo.normal_from_depth(g.pose.to(DEV), g.depth.to(DEV), g.intr.to(DEV), g.H, g.W)
    assert (n.cpu() - g.normal).abs().max() <= 2e-4      # unit normals from cancelling central differences
    zn, zf = compute_surfelinfo.depth_guided_range(g.depth.to(DEV), *synth.BG_RANGE)
    assert torch.equal(zn.cpu(), g.guided_near) and torch.equal(zf.cpu(), g.guided_far)


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sample_depth(torch.zeros(1, 4), torch.ones(1, 4), 8, stratified=False)


def test_view_matrices_kernel_matches_torch_to_the_last_bits():
    """tp_view_matrices (one launch; the bf16 path's per-view constants) against torch's inverse / Pose.invert: K^-1 is the
    fp64 cofactor formula rounded once, so it agrees with the LU-based fp32 inverse to a few ulp; so does pose^-1."""
    from texpose_b200 import camera
    B = 7
    pose = synth.poses(list(range(B))).to(DEV)
    intr = synth.intrinsics(B).to(DEV).clone()
    intr[3:, :2] *= 0.37
    k0, p0 = camera.view_matrices(pose, intr)
    k1, p1 = camera.view_matrices(pose, intr, one_launch=True)
    assert torch.equal(p0[..., :3], p1[..., :3])                                      # R^T: exact
    assert ((p0[..., 3] - p1[..., 3]).abs() <= 4 * torch.finfo(torch.float32).eps * 10.0).all()      # -(R^T t): |t| ~ 8
    exact = torch.linalg.inv(intr.double())
    assert (k1.double() - exact).abs().max() <= (k0.double() - exact).abs().max() + 1e-12      # at least as close to the exact inverse
    assert ((k1 - k0).abs() <= 4 * torch.finfo(torch.float32).eps * k0.abs().clamp(min=1e-3)).all()
