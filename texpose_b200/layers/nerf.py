"""Drop-in for layers/nerf.py: the plain NeRF (trainable trunk + rgb head, single-chain compositing)."""
from __future__ import annotations

import torch

from .. import ops
from . import _common
from ._mlp import MLPConfig, NerfMLP, run_mlp


class NeRF(torch.nn.Module):

    def __init__(self, opt):
        super().__init__()
        # layers/nerf.py:15-48
        d3 = 3 + 6 * opt.arch.posenc.L_3D if opt.arch.posenc else 3
        dview = (3 + 6 * opt.arch.posenc.L_view if opt.arch.posenc else 3) if opt.nerf.view_dep else 0
        tf = opt.arch.tf_init
        self.mlp_feat = torch.nn.ModuleList()
        dims = _common.layer_dims(opt.arch.layers_feat)
        for li, (k_in, k_out) in enumerate(dims):
            k_in = d3 if li == 0 else k_in
            k_in = k_in + d3 if li in opt.arch.skip else k_in
            last = li == len(dims) - 1
            lin = torch.nn.Linear(k_in, k_out + 1 if last else k_out)
            if tf:
                _common.tf_init_(lin, out="first" if last else None)
            self.mlp_feat.append(lin)
        self.mlp_rgb = torch.nn.ModuleList()
        dims = _common.layer_dims(opt.arch.layers_rgb)
        feat_dim = opt.arch.layers_feat[-1]
        for li, (k_in, k_out) in enumerate(dims):
            if li == 0:
                k_in = feat_dim + dview + 3
            lin = torch.nn.Linear(k_in, k_out)
            if tf:
                _common.tf_init_(lin, out="all" if li == len(dims) - 1 else None)
            self.mlp_rgb.append(lin)
        if opt.c2f is not None:
            self.progress = torch.nn.Parameter(torch.tensor(0.))

    def _config(self, opt, mode) -> MLPConfig:
        if opt.arch.density_activ != "softplus":
            raise NotImplementedError("texpose_b200 implements density_activ: softplus (the reference yaml setting)")
        if opt.nerf.density_noise_reg and mode == "train":
            raise NotImplementedError("density_noise_reg is null in every reference yaml; not implemented")
        if opt.c2f is not None:
            raise NotImplementedError("coarse-to-fine windowing is off (c2f null) in nerf_lm_env.yaml; not implemented")
        return MLPConfig(L_3D=opt.arch.posenc.L_3D, L_view=opt.arch.posenc.L_view, skip=tuple(opt.arch.skip),
                         view_dep=bool(opt.nerf.view_dep), n_feat=len(self.mlp_feat), n_rgb=len(self.mlp_rgb), n_trans=0,
                         precision="fp32", save_for_backward=torch.is_grad_enabled())

    def forward(self, opt, points_3D, ray_unit=None, mode=None):
        """layers/nerf.py:61-99 -> rgb [B,HW,N,3], density [B,HW,N]."""
        cfg = self._config(opt, mode)
        geom = _common.point_geometry(cfg, points_3D, ray_unit)
        return run_mlp(cfg, geom, None, None, *_common.flat_params(self.mlp_feat, self.mlp_rgb))

    def forward_samples(self, opt, center, ray, depth_samples, mode=None):
        """layers/nerf.py:101-115."""
        cfg = self._config(opt, mode)
        geom = _common.ray_geometry(cfg, center, ray, depth_samples)
        return run_mlp(cfg, geom, None, None, *_common.flat_params(self.mlp_feat, self.mlp_rgb))

    @staticmethod
    def composite(opt, ray, rgb_samples, density_samples, depth_samples):
        """layers/nerf.py:117-136 -> rgb, depth, opacity, prob."""
        bg = opt.data.bgcolor if opt.nerf.get("setbg_opaque") else None
        return ops.CompositePlain.apply(ray, rgb_samples, density_samples, depth_samples, bg)

    def positional_encoding(self, opt, x, L, c2f=False):
        if opt.c2f is not None and c2f:
            raise NotImplementedError("c2f windowing not implemented")
        flat = ops._f32(x.detach()).reshape(-1, 3)
        enc = ops.positional_encode(flat, L)
        return enc[:, 3:3 + 6 * L].reshape(*x.shape[:-1], 6 * L)
