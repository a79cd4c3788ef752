"""Mesh depth / NOCS / colour rasteriser (csrc/raster.cu, tp_mesh_render) against the CPU restatement
(oracle/mesh_oracle.py -- parity with pytorch3d itself is unpinned, the package is absent), its size-independent properties at
the 480 x 640 frame, the drop-in tools.mvrenderer.MVRenderer, and the 'render' range chain of data/lm.py:352-356: rasterised
depth -> 0.8 / 1.2 sample bounds -> Graph.render (BASELINE config C5)."""
import numpy as np
import pytest
import torch

from oracle import mesh_oracle as M
from oracle import texpose_oracle as O
from texpose_b200 import compute_surfelinfo, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.tools import mvrenderer
from tests.conftest import layer_list

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _poses(n):
    rows = []
    rng = np.random.RandomState(5)
    for i in range(n):
        a = rng.randn(3)
        a /= np.linalg.norm(a)
        th = 0.4 + 0.5 * i
        Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = np.array([0.05 * i - 0.04, 0.03 - 0.02 * i, 1.2 + 0.3 * i])
        rows.append(np.concatenate([R, t[:, None]], axis=1).reshape(12))
    return np.stack(rows).astype(np.float32)


def test_rasteriser_equals_the_oracle_bit_for_bit_on_coverage_and_depth():
    H, W = 48, 64
    K = np.array([[110.0, 0, 30.5], [0, 108.0, 25.25], [0, 0, 1]], np.float32)
    v, f = M.icosphere(2, 0.25)
    v = (v * np.array([1.0, 0.7, 1.3], np.float32)).astype(np.float32)           # an ellipsoid: no symmetric ties
    rng = np.random.RandomState(1)
    colors = rng.rand(v.shape[0], 3).astype(np.float32)
    poses = _poses(3)
    vt, ft = torch.from_numpy(v).to(DEV), torch.from_numpy(f).to(DEV)
    Kt = torch.from_numpy(K)[None].expand(3, 3, 3).contiguous().to(DEV)
    for name, attr in (("color", colors), ("nocs", M.nocs_coordinates(v))):
        out, depth, p2f = mvrenderer.render_mesh(vt, ft, torch.from_numpy(attr).to(DEV), torch.from_numpy(poses).to(DEV), Kt, H, W,
                                                 want_faces=True)
        for b in range(3):
            ro, rd, rf = M.render(v, f, attr, poses[b], K, H, W)
            assert (rf >= 0).sum() > 150
            assert np.array_equal(p2f[b].cpu().numpy(), rf), (name, b)
            assert np.array_equal(depth[b].cpu().numpy(), rd), (name, b)           # same arithmetic, one rounding per operation
            assert np.abs(out[b].cpu().numpy() - ro).max() <= 1e-6, (name, b)


def test_full_frame_properties_and_determinism():
    H, W = 480, 640
    Kt = synth.intrinsics(1).to(DEV)
    v, f = M.icosphere(5, 0.1)                                                     # 20 480 faces, ~1-2 pixels each
    vt, ft = torch.from_numpy(v).to(DEV), torch.from_numpy(f).to(DEV)
    nocs = mvrenderer.nocs_coordinates(vt)
    pose = torch.tensor([[1, 0, 0, 0.02, 0, 1, 0, -0.01, 0, 0, 1, 0.9]], device=DEV)
    out, depth, p2f = mvrenderer.render_mesh(vt, ft, nocs, pose, Kt, H, W, want_faces=True)
    out2, depth2, p2f2 = mvrenderer.render_mesh(vt, ft, nocs, pose, Kt, H, W, want_faces=True)
    assert torch.equal(out, out2) and torch.equal(depth, depth2) and torch.equal(p2f, p2f2)
    fx, fy, cx, cy = [float(Kt[0][i]) for i in ((0, 0), (1, 1), (0, 2), (1, 2))]
    cu, cv = cx + fx * 0.02 / 0.9, cy + fy * -0.01 / 0.9
    rr, cc = torch.meshgrid(torch.arange(H, device=DEV), torch.arange(W, device=DEV), indexing="ij")
    rad = torch.hypot((cc + 0.5 - cu) / fx, (rr + 0.5 - cv) / fy)                 # normalised image radius
    r_sil = 0.1 / (0.9 ** 2 - 0.1 ** 2) ** 0.5
    hit = p2f[0] >= 0
    assert bool(hit[rad < r_sil * 0.985].all()) and not bool(hit[rad > r_sil * 1.01].any())
    assert bool((depth[0][~hit] == -1).all()) and bool((out[0][:, ~hit] == 0).all())
    d = depth[0][hit]
    assert 0.7995 <= float(d.min()) <= 0.801 and float(d.max()) < 0.9             # front of the sphere at t_z - r
    assert float(out.min()) >= 0 and float(out.max()) <= 1
    # every visible face looks at the camera: its object-space z (NOCS blue) lies in the near half
    assert float(out[0, 2][hit].max()) < 0.56
    # a batch renders each view as the single-view call does
    pose3 = torch.from_numpy(_poses(3)).to(DEV)
    pose3[:, 11] = 0.9
    ob, db, _ = mvrenderer.render_mesh(vt, ft, nocs, pose3, Kt.expand(3, 3, 3).contiguous(), H, W)
    for b in range(3):
        o1, d1, _ = mvrenderer.render_mesh(vt, ft, nocs, pose3[b:b + 1], Kt, H, W)
        assert torch.equal(ob[b], o1[0]) and torch.equal(db[b], d1[0])


def test_mvrenderer_drop_in_and_surfel_normals():
    """MVRenderer(mesh, H, W, B)(pose, K, mode='color' | 'nocs') as compute_surfelinfo.py:114-116, then normal_from_depth."""
    H, W = 120, 160
    v, f = M.icosphere(4, 100.0)                                                   # millimetres, as the reference's CAD models
    rng = np.random.RandomState(2)
    mesh = mvrenderer.Mesh(torch.from_numpy(v), torch.from_numpy(f.astype(np.int64)), torch.from_numpy(rng.rand(len(v), 3).astype(np.float32))).to(DEV)
    r = mvrenderer.MVRenderer(mesh, H, W, 1, None, mode="complex")
    K = torch.tensor([[[150.0, 0, 80], [0, 150.0, 60], [0, 0, 1]]], device=DEV)
    pose = torch.tensor([[[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 900.0]]], device=DEV)
    rgb, depth = r(pose, K, mode="color", return_depth=True)
    nocs, depth2 = r(pose, K, mode="nocs", return_depth=True)
    assert rgb.shape == (1, 3, H, W) and depth.shape == (1, H, W) and torch.equal(depth, depth2)
    assert r(pose, K, mode="nocs", return_depth=False).shape == (1, 3, H, W)
    with pytest.raises(NotImplementedError):
        r(pose, K, mode="mask")
    feat, _ = r(pose, K, mode="feature")                                           # vertex features: the same attribute path
    assert torch.equal(feat, rgb)
    hit = depth[0] > 0
    assert 840 < int(hit.sum()) < 920 and abs(float(depth[0, 60, 80]) - 800.0) < 0.5
    n = compute_surfelinfo.normal_from_depth(pose, depth.clamp(min=0) / 1000.0, K, h=H, w=W)
    centre = n[0, :, 60, 80]
    assert float(centre.abs()[2]) > 0.99                                           # the sphere faces the camera at the image centre
    # the Pose wrapper layout of the reference (nine rotation entries, then t) gives the same render
    flat = torch.cat([pose[:, :, :3].reshape(1, 9), pose[:, :, 3]], dim=-1)
    rgb_b, _ = r(flat, K, mode="color")
    assert torch.equal(rgb, rgb_b)


def test_render_range_from_rasterised_depth_feeds_graph_render():
    """BASELINE C5: depth of the CAD mesh -> z_near / z_far = 0.8 / 1.2 x depth (data/lm.py:352-356) -> Graph.render, against the
    oracle rendering the same rays over the same bounds (fp32 mode, 1e-4)."""
    from texpose_b200.model.nerf_adapt_st_gan import Graph
    H, W, N = 48, 64, 32
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.nerf.sample_stratified = False
    opt.b200 = AttrDict(mlp="fp32")
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=2).to(DEV)
    pose, intr = synth.poses([0]), synth.intrinsics(1).clone()
    intr[:, :2] *= 0.1
    v, f = M.icosphere(3, 0.85)
    rows = pose.reshape(1, 12).to(DEV)
    _, depth, _ = mvrenderer.render_mesh(torch.from_numpy(v).to(DEV), torch.from_numpy(f).to(DEV), None, rows, intr.to(DEV), H, W)
    _, rd, _ = M.render(v, f, np.zeros((len(v), 1), np.float32), pose.reshape(12).numpy(), intr[0].numpy(), H, W)
    assert np.array_equal(depth[0].cpu().numpy(), rd) and (rd > 0).sum() > 60
    zn, zf = compute_surfelinfo.depth_guided_range(depth.clamp(min=0).reshape(1, H * W), *synth.BG_RANGE)
    ozn, ozf = O.depth_guided_range(torch.from_numpy(rd).clamp(min=0).reshape(1, H * W), *synth.BG_RANGE)
    assert torch.equal(zn.cpu(), ozn) and torch.equal(zf.cpu(), ozf)
    obj = torch.from_numpy(rd > 0).reshape(-1).nonzero()[:, 0]
    with torch.no_grad():
        out = g.render(opt, pose.to(DEV), intr=intr.to(DEV), ray_idx=obj[None].to(DEV), depth_range=(zn[:, :, None], zf[:, :, None]),
                       sample_idx=None, mode="val")
    c, r = O.get_center_and_ray(pose, intr, H, W)
    ref = O.render_stl(O.gather_rays(c, obj[None]), O.gather_rays(r, obj[None]), ozn[:, obj], ozf[:, obj], None, N,
                       g.latent_vars_trans.weight[0][None].cpu(), g.latent_vars_light.weight[0][None].cpu(),
                       [(w.cpu(), b.cpu()) for w, b in layer_list(g.nerf.mlp_feat)], [(w.cpu(), b.cpu()) for w, b in layer_list(g.nerf.mlp_rgb)],
                       [(w.cpu(), b.cpu()) for w, b in layer_list(g.nerf.mlp_trans)])
    for k in ("rgb", "depth", "opacity", "uncert"):
        assert (out[k].cpu() - ref[k]).abs().max() <= 1e-4, k


def test_argument_errors():
    from texpose_b200 import _C, ops
    lib = _C.load()
    z = torch.zeros(64, device=DEV)
    fi = torch.zeros(3, dtype=torch.int32, device=DEV)
    ws = torch.zeros(lib.tp_mesh_render_workspace(1, 3, 8, 8), dtype=torch.uint8, device=DEV)
    import ctypes
    args = lambda **kw: [kw.get("verts", ops._p(z)), 3, ops._p(fi), 1, ops._p(z), kw.get("C", 3), ops._p(z), ops._p(z), 1, 8, 8,
                         ctypes.c_float(kw.get("sigma", 1e-4)), ops._p(z), ops._p(z), None, ops._p(ws), kw.get("ws", ws.numel()), None]
    assert lib.tp_mesh_render(*args(verts=None)) == -1
    assert lib.tp_mesh_render(*args(C=9)) == -2
    assert lib.tp_mesh_render(*args(sigma=0.0)) == -2
    assert lib.tp_mesh_render(*args(ws=16)) == -5


def test_empty_mesh_and_faces_behind_the_camera_render_background():
    v, f = M.icosphere(1, 0.2)
    vt, ft = torch.from_numpy(v).to(DEV), torch.from_numpy(f).to(DEV)
    Kt = torch.tensor([[[100.0, 0, 32], [0, 100.0, 24], [0, 0, 1]]], device=DEV)
    front = torch.tensor([[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 2.0]], device=DEV)
    behind = torch.tensor([[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, -2.0]], device=DEV)
    attr = torch.ones(v.shape[0], 3, device=DEV)
    out, depth, p2f = mvrenderer.render_mesh(vt, ft[:0], attr, front, Kt, 48, 64, want_faces=True)      # no faces at all
    assert bool((depth == -1).all()) and bool((p2f == -1).all()) and float(out.abs().max()) == 0
    out, depth, p2f = mvrenderer.render_mesh(vt, ft, attr, behind, Kt, 48, 64, want_faces=True)
    assert bool((depth == -1).all()) and bool((p2f == -1).all())
    out, depth, _ = mvrenderer.render_mesh(vt, ft, attr, front, Kt, 48, 64)
    assert int((depth > 0).sum()) > 200 and abs(float(out[0, 0][depth[0] > 0].min()) - 1) < 1e-6
