"""Option tree (`opt`) mirror for the render hot path.

The reference threads an ``EasyDict opt`` through every call (options.py:17-141,
options/nerf_lm_adapt_gan.yaml).  Only the keys the hot path reads are restated
here; any attribute-style dict (including the reference's own EasyDict) is
accepted by the kernels' host wrappers.
"""
from __future__ import annotations

import copy


class AttrDict(dict):
    """Minimal attribute dict (stand-in for easydict.EasyDict, which is not installed)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return AttrDict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(AttrDict._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # mirror EasyDict behaviour
            raise AttributeError(k) from e

    def update(self, other=None, **kw):
        for k, v in dict(other or {}, **kw).items():
            self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def adapt_gan_opt(H=480, W=640, sample_intvs=64, device="cpu", **over):
    """Options of options/nerf_lm_adapt_gan.yaml:3-40,66-68,114-118 that the render path reads."""
    opt = AttrDict(
        model="nerf_adapt_st_gan",
        device=device, H=H, W=W,
        c2f=dict(range=None, start=None),
        arch=dict(
            layers_feat=[None, 256, 256, 256, 256, 256, 256, 256, 256],
            layers_rgb=[None, 256, 256, 256, 3],
            layers_trans=[None, 256, 256, 256, 5],
            skip=[4],
            posenc=dict(L_3D=10, L_view=4),
            density_activ="softplus",
            tf_init=True,
        ),
        nerf=dict(
            view_dep=True,
            depth=dict(param="metric", range=[0, 3], scale=10, range_source="box",
                       box_mask=False, box_source="pred_box_init_calib"),
            sample_intvs=sample_intvs,
            sample_stratified=True,
            rand_rays=2048,
            density_noise_reg=None,
            mask_obj=True,
            N_latent=32, N_latent_trans=16, N_latent_light=48,
            min_uncert=0.05,
        ),
        camera=dict(model="perspective", ndc=False),
        data=dict(image_size=[H, W], pose_source="predicted", bgcolor=None),
        render=dict(N_candidate=3, transient="zero"),
        loss_weight=dict(render=0, uncert=0, trans_reg=-2),
        batch_size=8, patch_size=16,
    )
    for k, v in over.items():
        node = opt
        parts = k.split("__")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return opt


def env_opt(H=128, W=128, sample_intvs=64, device="cpu"):
    """Options of options/nerf_lm_env.yaml (plain layers/nerf.py NeRF: c2f null, 128-wide rgb head)."""
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=sample_intvs, device=device)
    opt.model = "nerf_pretrain_env"
    opt.c2f = None
    opt.arch.layers_rgb = [None, 128, 3]
    opt.arch.layers_trans = None
    opt.nerf.setbg_opaque = None
    return opt
