"""Size-independent properties of the oracle's restatements (beyond the fixtures of the real reference): what has to hold for ANY
input of the path.  The GPU suite checks the same properties on the CUDA path at BASELINE's full sizes (tests/test_gpu_fullsize.py);
here they guard the checker itself.  CPU only, seeded."""
import torch

from oracle import texpose_oracle as O
from texpose_b200 import synth


def _rays(n=500, seed=0):
    pose, intr = synth.poses([seed]), synth.intrinsics(1)
    c, r = O.get_center_and_ray(pose, intr, 480, 640)
    idx = torch.randperm(480 * 640, generator=torch.Generator().manual_seed(seed))[:n][None]
    return O.gather_rays(c, idx), O.gather_rays(r, idx)


def test_rays_have_unit_camera_depth_and_a_common_centre():
    """camera.py:292-314: every ray leaves the camera centre and has camera-z component 1 (so a sample's depth is its camera z)."""
    pose, intr = synth.poses([3]), synth.intrinsics(1)
    c, r = O.get_center_and_ray(pose, intr, 480, 640)
    assert (c - c[:, :1]).abs().max() == 0
    R = pose[0, :, :3]
    z = (r[0] @ R.T)[:, 2]                                  # back to the camera frame
    assert (z - 1).abs().max() <= 2e-6
    assert (c[0, 0] + R.T @ pose[0, :, 3]).abs().max() <= 1e-6      # centre = -R^T t


def test_aabb_hits_lie_on_the_box_and_misses_stay_outside():
    c, r = _rays(4000, seed=1)
    lo, hi = synth.padded_aabb()
    tn, tf, valid = O.aabb_ray_intersection(lo, hi, c, r)
    assert valid.any() and (~valid).any()
    v = valid[0]
    assert (tf[0, v] > tn[0, v]).all() and (tf[0, v] > 0).all()
    for t in (tn, tf):
        p = c[0, v] + r[0, v] * t[0, v, None]
        inside = ((p >= lo.view(1, 3) - 1e-4) & (p <= hi.view(1, 3) + 1e-4)).all(dim=1)
        on_face = torch.minimum((p - lo.view(1, 3)).abs(), (p - hi.view(1, 3)).abs()).min(dim=1).values <= 1e-4
        assert inside.all() and on_face.all()
    mid = c[0, ~v] + r[0, ~v] * 8.0                          # a missing ray's point at the object's distance is outside the box
    assert (~((mid >= lo.view(1, 3)) & (mid <= hi.view(1, 3))).all(dim=1)).all()


def test_sample_depths_are_sorted_stratified_and_inside_their_bins():
    g = torch.Generator().manual_seed(2)
    zn = torch.rand(2, 50, generator=g) * 5 + 1
    zf = zn + torch.rand(2, 50, generator=g) * 3 + 0.1
    for n in (32, 64, 128):
        rand = torch.rand(2, 50, n, 1, generator=g)
        d = O.sample_depth(zn, zf, n, rand)[..., 0]
        assert (d[..., 1:] > d[..., :-1]).all()
        width = ((zf - zn) / n)[..., None]
        k = torch.arange(n).view(1, 1, n)
        assert (d >= zn[..., None] + k * width - 1e-5).all() and (d <= zn[..., None] + (k + 1) * width + 1e-5).all()
        mid = O.sample_depth(zn, zf, n, None)[..., 0]
        assert ((mid - (zn[..., None] + (k + 0.5) * width)).abs() <= 1e-5).all()


def test_compositing_is_a_partition_of_unity_linear_in_colour_and_opaque_at_the_last_sample():
    """layers/nerf_static_transient_light.py:168-212: the weights T*alpha of a chain sum to 1 - T_end; the 1e10 tail makes the last
    sample opaque wherever its density is positive, so every chain's opacity is 1 there; outputs are linear in the colours."""
    g = torch.Generator().manual_seed(3)
    B, R, N = 2, 40, 48
    ray = torch.randn(B, R, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])
    depth = (torch.rand(B, R, N, 1, generator=g) + torch.arange(N).view(1, 1, N, 1)) / N * 2 + 7
    rgb = torch.rand(B, R, N, 3, 2, generator=g)
    dens = torch.rand(B, R, N, 2, generator=g) * 3 + 1e-3
    unc = torch.rand(B, R, N, 1, generator=g)
    out = O.composite_stl(ray, rgb, dens, depth, unc, 0.05)
    names = ["rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient", "prob", "uncert",
             "alpha_static", "alpha_transient"]
    o = dict(zip(names, out))
    for k in ("opacity", "opacity_static", "opacity_transient"):
        assert (o[k] - 1).abs().max() <= 1e-5, k
    assert (o["prob"].sum(dim=2) - o["opacity"]).abs().max() <= 1e-6
    assert (o["prob"] >= 0).all() and (o["alpha_static"] >= 0).all() and (o["alpha_static"] <= 1).all()
    assert (o["depth"] >= depth.min() - 1e-4).all() and (o["depth"] <= depth.max() + 1e-4).all()      # a convex combination of the depths
    assert (o["uncert"] >= 0.05).all()
    rgb2 = torch.rand(B, R, N, 3, 2, generator=g)
    a, b = 0.3, 1.7
    mix = O.composite_stl(ray, a * rgb + b * rgb2, dens, depth, unc, 0.05)
    other = O.composite_stl(ray, rgb2, dens, depth, unc, 0.05)
    for i in range(3):
        assert (mix[i] - (a * out[i] + b * other[i])).abs().max() <= 1e-5
    # the plain chain is the static chain of the same densities
    p = O.composite_plain(ray, rgb[..., 0], dens[..., 0], depth)
    assert (p[0] - o["rgb_static"]).abs().max() <= 1e-6 and (p[1] - o["depth"]).abs().max() <= 1e-5
    assert (p[2] - o["opacity_static"]).abs().max() <= 1e-6


def test_positional_encoding_layout_and_zero_density_renders_nothing():
    x = torch.tensor([[0.25, -0.5, 1.0]])
    enc = O.positional_encoding(x, 4)                       # per coordinate [sin f0..f3, cos f0..f3], f_k = 2^k pi (layers/...light.py:217-234)
    assert enc.shape == (1, 24)
    f = (2.0 ** torch.arange(4)) * torch.pi
    want = torch.cat([torch.cat([torch.sin(x[0, i] * f), torch.cos(x[0, i] * f)]) for i in range(3)])
    assert (enc[0] - want).abs().max() <= 1e-6
    ray = torch.tensor([[[0.0, 0.0, 1.0]]])
    depth = torch.linspace(7, 9, 8).view(1, 1, 8, 1)
    out = O.composite_plain(ray, torch.rand(1, 1, 8, 3), torch.zeros(1, 1, 8), depth)
    assert out[0].abs().max() == 0 and out[2].abs().max() == 0
