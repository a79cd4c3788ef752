"""CPU checks of the mesh rasteriser restatement (oracle/mesh_oracle.py; SURVEY 8 f4, parity unpinned: pytorch3d is absent)
against closed forms: a fronto-parallel quad (exact coverage, constant depth, affine attribute interpolation), occlusion order,
a sphere's silhouette and depth, and the NOCS vertex attribute of tools/mvrenderer.py:695-722."""
import numpy as np

from oracle import mesh_oracle as M

K = np.array([[100.0, 0, 32.0], [0, 100.0, 24.0], [0, 0, 1]], np.float32)
H, W = 48, 64
EYE = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)      # [R | t] rows


def _quad(x0, x1, y0, y1, z):
    v = np.array([[x0, y0, z], [x1, y0, z], [x1, y1, z], [x0, y1, z]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    return v, f


def test_fronto_parallel_quad_coverage_depth_and_interpolation():
    v, f = _quad(-0.52, 0.43, -0.31, 0.37, 5.0)
    attr = np.stack([v[:, 0], v[:, 1], np.ones(4, np.float32)], axis=1)      # attribute = the vertex' own x, y: affine in the image
    out, depth, p2f = M.render(v, f, attr, EYE, K, H, W)
    u0, u1, v0, v1 = 32 + 100 * -0.52 / 5, 32 + 100 * 0.43 / 5, 24 + 100 * -0.31 / 5, 24 + 100 * 0.37 / 5
    cols = [c for c in range(W) if u0 < c + 0.5 < u1]
    rows = [r for r in range(H) if v0 < r + 0.5 < v1]
    want = np.zeros((H, W), bool)
    want[np.ix_(rows, cols)] = True
    assert np.array_equal(p2f >= 0, want)
    assert np.abs(depth[want] - 5.0).max() < 2e-6 and np.all(depth[~want] == -1.0)      # (w0 + w1 + w2 = 1 up to rounding)
    rr, cc = np.nonzero(want)
    x_true, y_true = (cc + 0.5 - 32) * 5 / 100, (rr + 0.5 - 24) * 5 / 100
    assert np.abs(out[0][want] - x_true).max() < 1e-5 and np.abs(out[1][want] - y_true).max() < 1e-5
    assert np.abs(out[2][want] - 1).max() < 1e-6 and np.all(out[:, ~want] == 0)


def test_nearest_face_wins_and_ties_keep_the_lower_index():
    va, fa = _quad(-0.5, 0.5, -0.5, 0.5, 6.0)
    vb, fb = _quad(-0.2, 0.2, -0.2, 0.2, 4.0)
    v, f = np.concatenate([va, vb]), np.concatenate([fa, fb + 4])
    attr = np.concatenate([np.tile([1.0, 0, 0], (4, 1)), np.tile([0, 1.0, 0], (4, 1))]).astype(np.float32)
    out, depth, p2f = M.render(v, f, attr, EYE, K, H, W)
    assert abs(depth[24, 32] - 4.0) < 2e-6 and p2f[24, 32] >= 2 and out[1, 24, 32] > 0.99999
    assert abs(depth[24, 38] - 6.0) < 2e-6 and p2f[24, 38] < 2 and depth[24, 41] == -1
    v2, f2 = np.concatenate([va, va]), np.concatenate([fa, fa + 4])           # the same quad twice: equal depths
    _, _, p2 = M.render(v2, f2, np.ones((8, 1), np.float32), EYE, K, H, W)
    assert p2.max() <= 1


def test_sphere_silhouette_depth_and_pose():
    v, f = M.icosphere(3, 0.2)
    pose = np.array([1, 0, 0, 0.05, 0, 1, 0, -0.03, 0, 0, 1, 1.5], np.float32)
    out, depth, p2f = M.render(v, f, M.nocs_coordinates(v), pose, K, H, W)
    cu, cv = 32 + 100 * 0.05 / 1.5, 24 + 100 * -0.03 / 1.5
    rr, cc = np.mgrid[0:H, 0:W]
    rad = np.hypot(cc + 0.5 - cu, rr + 0.5 - cv)
    r_sil = 100 * 0.2 / np.sqrt(1.5 ** 2 - 0.2 ** 2)                           # tangent-cone radius in pixels
    assert np.all(p2f[rad < r_sil - 1.0] >= 0) and np.all(p2f[rad > r_sil + 0.5] < 0)
    assert abs(depth[int(cv), int(cu)] - 1.3) < 2e-3                           # front of the sphere
    assert depth[p2f >= 0].max() < 1.5 and out.min() >= 0 and out.max() <= 1
    assert out[2][p2f >= 0].max() < 0.55                                       # only the camera-facing half (low object z) is visible


def test_nocs_coordinates_formula():
    v = np.array([[0, 0, 0], [2, 0, 0], [0, 4, 0], [0, 0, 8]], np.float32)
    n = M.nocs_coordinates(v)
    ct = v.mean(0)
    want = ((v - ct) / np.abs(v - ct).max(0) + 1) / 2
    assert np.allclose(n, want) and n.min() >= 0 and n.max() <= 1


def test_pixel_centres_are_half_integers_of_the_opencv_projection():
    # x_ndc of pixel column c equals -((c + 0.5) - W/2) / (min(H, W) / 2), likewise for rows
    s = min(H, W) / 2
    for c in (0, 13, W - 1):
        assert abs(M.pix_to_ndc(W - 1 - c, W, H) + ((c + 0.5) - W / 2) / s) < 1e-6
    for r in (0, 7, H - 1):
        assert abs(M.pix_to_ndc(H - 1 - r, H, W) + ((r + 0.5) - H / 2) / s) < 1e-6


def test_drop_in_host_helpers_on_cpu():
    """tools.mvrenderer host side without a GPU: NOCS vertex attribute == the oracle's, every pose layout the reference passes
    (Pose wrapper with .R / .t, [B,3,4], [B,4,4], the wrapper's flat [R(9) | t(3)] data) maps to the same [R | t] rows, the Mesh
    holder answers what MVRenderer reads from a pytorch3d Meshes, and a CPU mesh is refused loudly (no CPU path)."""
    import pytest
    import torch
    from texpose_b200.tools import mvrenderer as mv
    v, f = M.icosphere(2, 0.3)
    vt = torch.from_numpy(v) * torch.tensor([1.0, 2.0, 0.5])
    assert np.abs(mv.nocs_coordinates(vt).numpy() - M.nocs_coordinates(vt.numpy())).max() < 1e-6
    R = torch.tensor([[[0.0, -1, 0], [1, 0, 0], [0, 0, 1]]])
    t = torch.tensor([[0.1, 0.2, 3.0]])
    want = torch.cat([R, t[:, :, None]], dim=-1).reshape(1, 12)

    class PoseLike:          # the reference's Pose wrapper exposes .R and .t (tools/mvrenderer.py:495-503)
        def __init__(self, R, t):
            self.R, self.t = R, t

    T44 = torch.eye(4)[None].clone()
    T44[:, :3, :3], T44[:, :3, 3] = R, t
    flat = torch.cat([R.reshape(1, 9), t], dim=-1)
    for pose in (PoseLike(R, t), torch.cat([R, t[:, :, None]], dim=-1), T44, flat):
        assert torch.equal(mv._pose_rows(pose), want)
    mesh = mv.Mesh(vt, torch.from_numpy(f.astype(np.int64)), torch.rand(len(v), 3))
    assert mesh.verts_packed().shape == (len(v), 3) and mesh.faces_packed().shape == (len(f), 3)
    assert mesh.textures.verts_features_packed().shape == (len(v), 3) and mesh.extend(4) is mesh
    r = mv.MVRenderer(mesh, 48, 64, 1, None)
    assert set(r.attrs) == {"nocs", "color"} and r.faces.dtype == torch.int32
    with pytest.raises(RuntimeError):
        r(torch.cat([R, t[:, :, None]], dim=-1), torch.eye(3)[None], mode="nocs")
    with pytest.raises(NotImplementedError):
        r(torch.cat([R, t[:, :, None]], dim=-1), torch.eye(3)[None], mode="mask")
