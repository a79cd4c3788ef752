"""Drop-in for tools/mvrenderer.py (MVRenderer, :33-178) as compute_surfelinfo.py:114-116 uses it: depth + vertex-colour /
NOCS maps of a CAD mesh under a batch of poses, on the rasteriser of csrc/raster.cu (tp_mesh_render) instead of pytorch3d.

pytorch3d is neither vendored nor pinned by the reference (README.md:16) and absent here, so the kernel restates its
published rasterisation rules; parity with the real package is UNPINNED.  Modes 'color' and 'nocs' (the
two compute_surfelinfo renders) and 'feature' (vertex features of up to 8 channels) are implemented; 'mask' (70 faces per pixel
soft silhouette) and 'normal' (tangent-space normal maps) are not.
"""
from __future__ import annotations

import torch

from .. import _C, ops


class Mesh:
    """The three things MVRenderer reads from a pytorch3d `Meshes` with `TexturesVertex`: packed vertices [V,3], packed
    faces [F,3] and per-vertex colours [V,3] (compute_surfelinfo.py:84-90).  A real `Meshes` object works as well."""

    def __init__(self, verts, faces, colors=None):
        self._verts = verts.reshape(-1, 3).float()
        self._faces = faces.reshape(-1, 3)
        self._colors = None if colors is None else colors.reshape(self._verts.shape[0], -1).float()
        self.textures = self
        self.device = self._verts.device

    def verts_packed(self):
        return self._verts

    def faces_packed(self):
        return self._faces

    def verts_features_packed(self):
        return self._colors

    def extend(self, n):
        return self

    def to(self, device):
        return Mesh(self._verts.to(device), self._faces.to(device), None if self._colors is None else self._colors.to(device))


def nocs_coordinates(verts):
    """Vertex attribute of SoftPhongNOCSShader (mvrenderer.py:695-722): ((v - mean) / max|v - mean| + 1) / 2 per axis."""
    ct = verts.mean(dim=0, keepdim=True)
    d = verts - ct
    return (d / d.abs().max(dim=0, keepdim=True)[0] + 1) / 2


def _pose_rows(pose):
    """[B,12] = [R | t] rows from the reference's `Pose` wrapper (.R, .t), a [B,3,4] / [B,4,4] matrix or a [B,12] tensor."""
    if hasattr(pose, "R") and hasattr(pose, "t") and not torch.is_tensor(pose):
        R, t = pose.R, pose.t
        return torch.cat([R.reshape(-1, 3, 3), t.reshape(-1, 3, 1)], dim=-1).reshape(-1, 12)
    pose = torch.as_tensor(pose)
    if pose.shape[-2:] == (4, 4):
        pose = pose[..., :3, :]
    if pose.shape[-2:] == (3, 4):
        return pose.reshape(-1, 12)
    if pose.shape[-1] == 12:      # Pose._data layout: nine rotation entries, then t (mvrenderer.py:451-464)
        R, t = pose[..., :9].reshape(-1, 3, 3), pose[..., 9:].reshape(-1, 3, 1)
        return torch.cat([R, t], dim=-1).reshape(-1, 12)
    raise ValueError(f"unsupported pose shape {tuple(pose.shape)}")


class MVRenderer(torch.nn.Module):
    def __init__(self, cad_mesh, height, width, batch_size, cam_K=None, mode="complex"):
        super().__init__()
        if mode not in ("simplified", "complex"):
            raise NotImplementedError(mode)
        mesh = cad_mesh[0] if isinstance(cad_mesh, (list, tuple)) else cad_mesh
        self.device = torch.device(mesh.device)
        self.height, self.width, self.batch_size = height, width, batch_size
        verts = mesh.verts_packed().float().to(self.device)
        faces = mesh.faces_packed().to(self.device)
        if hasattr(mesh, "num_verts_per_mesh"):      # a pytorch3d Meshes extended to the batch: keep the first copy
            n = int(mesh.num_verts_per_mesh()[0])
            f = int(mesh.num_faces_per_mesh()[0])
            verts, faces = verts[:n], faces[:f]
        self.verts = verts.contiguous()
        self.faces = faces.to(torch.int32).contiguous()
        colors = mesh.textures.verts_features_packed() if getattr(mesh, "textures", None) is not None else None
        self.attrs = {"nocs": nocs_coordinates(self.verts).contiguous()}
        if colors is not None:
            self.attrs["color"] = colors.float().to(self.device)[:self.verts.shape[0]].contiguous()      # [V, C]: colours or any vertex features
        if cam_K is None:
            cam_K = torch.eye(3)
        self.K = torch.as_tensor(cam_K, dtype=torch.float32).reshape(-1, 3, 3)[:1].to(self.device)
        self.sigma = 1e-4      # BlendParams(sigma=1e-4, gamma=1e-4) of every renderer (mvrenderer.py:85,108)

    def forward(self, pose, K=None, mode="feature", return_depth=True):
        """mvrenderer.py:152-178 -> rendered [B,3,H,W] (+ depth [B,H,W], -1 = background, from fragments.zbuf)."""
        if mode == "feature" and "color" in self.attrs and self.attrs["color"].shape[1] <= 8:
            mode = "color"      # SoftPhongFeatureShader (mvrenderer.py:929-956): the vertex features themselves, same blend
        if mode not in self.attrs:
            raise NotImplementedError(f"MVRenderer mode {mode!r}: texpose_b200 implements 'color' / 'feature' (vertex features, <= 8 "
                                      "channels) and 'nocs' (compute_surfelinfo.py:114-115 renders 'color' and 'nocs')")
        if self.device.type != "cuda":
            raise RuntimeError("texpose_b200 has no CPU path: MVRenderer needs a CUDA mesh")
        rows = _pose_rows(pose).float().to(self.device).contiguous()
        B = rows.shape[0]
        Kb = (self.K if K is None else torch.as_tensor(K, dtype=torch.float32).reshape(-1, 3, 3).to(self.device))
        Kb = Kb.expand(B, 3, 3).contiguous()
        out, depth, _ = render_mesh(self.verts, self.faces, self.attrs[mode], rows, Kb, self.height, self.width, self.sigma)
        return (out, depth) if return_depth else out


def render_mesh(verts, faces, attr, pose_rows, K, H, W, sigma=1e-4, want_faces=False):
    """tp_mesh_render -> (out [B,C,H,W] or None, depth [B,H,W], pix_to_face [B,H,W] int32 or None)."""
    dev = verts.device
    B, V, F = pose_rows.shape[0], verts.shape[0], faces.shape[0]
    C = 0 if attr is None else attr.shape[1]
    out = torch.empty(B, C, H, W, device=dev) if attr is not None else None
    depth = torch.empty(B, H, W, device=dev)
    p2f = torch.empty(B, H, W, dtype=torch.int32, device=dev) if want_faces else None
    ws = torch.empty(_C.load().tp_mesh_render_workspace(B, V, H, W), dtype=torch.uint8, device=dev)
    _C.call("tp_mesh_render", ops._p(verts), V, ops._p(faces), F, ops._p(attr), C, ops._p(pose_rows), ops._p(K), B, H, W,
            float(sigma), ops._p(out), ops._p(depth), ops._p(p2f), ops._p(ws), ws.numel(), ops._stream())
    return out, depth, p2f
