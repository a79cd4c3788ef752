"""Host side of the tensor-core backward (csrc/mlp_tc_bwd.cu): transposed weight image, kernel sequencing.

Gradient sinks are the reference's: mlp_rgb.*, mlp_trans.*, the two latents (model/nerf_adapt_st_gan.py:56-69,108-127);
the trunk is frozen (layers/nerf_static_transient_light.py:34).
"""
from __future__ import annotations

import ctypes

import torch

from . import _C, ops

FWD_SLOTS, DZ_SLOTS = 7, 6
# forward-save slots: 0 feat, 1-3 rgb h1..h3, 4-6 trans h1..h3; dz slots: rgb dz2,dz1,dz0, trans dz2,dz1,dz0
_BIG_PAIRS = {"rgb": [(0, 2), (1, 1), (2, 0)], "trans": [(3, 5), (4, 4), (5, 0)]}      # (dz slot, x slot) for layers 2,1,0


def _bwd_table(rgb_p, trans_p):
    rows = []
    for layers in (rgb_p, trans_p):
        W3, W2, W1 = layers[3][0], layers[2][0], layers[1][0]
        rows.append([W3.data_ptr(), W3.stride(0), 0, 256, 0, W3.shape[0], 256, 0, -1, 1])     # K=16 chunk: W3^T
        for W in (W2, W1):
            for c in range(0, 256, 32):
                rows.append([W.data_ptr(), W.stride(0), 0, 256, c, 32, 256, 0, -1, 1])
    return rows


def pack_bwd(holder, rgb_p, trans_p):
    params = [w for w, _ in (rgb_p[1:] + trans_p[1:])]
    key = tuple((p.data_ptr(), p._version) for p in params)
    cached = getattr(holder, "_packed_bwd", None) if holder is not None else None
    if cached is not None and cached[0] == key:
        return cached[1]
    lib = _C.load()
    table = _bwd_table(rgb_p, trans_p)
    assert len(table) == lib.tp_tc_bwd_num_chunks()
    dev = rgb_p[0][0].device
    desc = torch.tensor(table, dtype=torch.int64, device=dev)
    img = torch.empty(len(table) * lib.tp_tc_chunk_bytes(), dtype=torch.uint8, device=dev)
    _C.call("tp_tc_pack_weights", ops._p(desc), len(table), ops._p(img), ops._stream())
    if holder is not None:
        holder._packed_bwd = (key, img, desc)
    return img


def backward_chain(dz_rgb, dz_trans, S, packed_bwd, saved):
    dz_images = torch.empty(_C.load().tp_tc_dz_bytes(S), dtype=torch.uint8, device=dz_rgb.device)
    _C.call("tp_tc_backward_chain", ops._p(dz_rgb), ops._p(dz_trans), S, ops._p(packed_bwd), ops._p(saved),
            ops._p(dz_images), ops._stream())
    return dz_images


def reduce_partials(partial, splits, count, out=None):
    out = torch.empty(count, device=partial.device) if out is None else out
    _C.call("tp_reduce_partials", ops._p(partial), splits, count, ops._p(out), 0, ops._stream())
    return out


def dw_gemm(a_images, a_slot, a_nslots, b_images, b_slot, b_nslots, S, flags=0):
    """[256,256] fp32 = dz^T x over all samples (split over tiles, fixed-order reduce)."""
    grid = _C.load().tp_tc_dw_grid(S)
    partial = torch.empty(grid * 65536, device=a_images.device)
    _C.call("tp_tc_dw_gemm", ops._p(a_images), a_slot, a_nslots, ops._p(b_images), b_slot, b_nslots, S, ops._p(partial),
            partial.numel(), flags, ops._stream())
    return reduce_partials(partial, grid, 65536).view(256, 256)


def thin_dw(thin, images, slot, n_slots, S):
    """[M,256] = thin^T x for a thin fp32 operand [S,M] (M in 1,3,5) and an image slot x."""
    M = thin.shape[1]
    max_blocks = 148 * 4 + 8
    partial = torch.empty(max_blocks * M * 256, device=thin.device)
    nb = ctypes.c_int(0)
    _C.call("tp_tc_thin_dw", ops._p(thin), M, ops._p(images), slot, n_slots, S, ops._p(partial), partial.numel(),
            ctypes.byref(nb), ops._stream())
    return reduce_partials(partial, nb.value, M * 256).view(M, 256)


def unpack(images, slot, n_slots, S):
    out = torch.empty(S, 256, device=images.device)
    _C.call("tp_tc_unpack_images", ops._p(images), slot, n_slots, S, ops._p(out), ops._stream())
    return out


def heads_backward(cfg, sv, S, per_image, rgb_p, trans_p, g_rgb, g_density, g_uncert, need_lat_trans, need_lat_light):
    """Returns (rgb layer grads [(dW,db)]*4, trans layer grads, d_lat_trans, d_lat_light)."""
    dev = sv.rgb.device
    geom = sv.geom
    lt, ll = sv.lat
    B, R, N = geom["shape"]
    dz_rgb = torch.empty(S, 3, device=dev)
    dz_trans = torch.empty(S, 5, device=dev)
    _C.call("tp_stl_output_grad", ops._p(sv.rgb), ops._p(sv.density), ops._p(sv.uncert), ops._p(g_rgb), ops._p(g_density),
            ops._p(g_uncert), S, ops._p(dz_rgb), ops._p(dz_trans), None, ops._stream())
    packed = pack_bwd(cfg.packed, rgb_p, trans_p)
    dz = backward_chain(dz_rgb, dz_trans, S, packed, sv.images)
    ones = torch.ones(S, 1, device=dev)
    out = {}
    for head, layers, dz3, h3_slot in (("rgb", rgb_p, dz_rgb, 3), ("trans", trans_p, dz_trans, 6)):
        grads = [None] * 4
        # output layer: thin weight gradient + bias
        grads[3] = (thin_dw(dz3, sv.images, h3_slot, FWD_SLOTS, S), ops.group_colsum(dz3, S, S).view(-1))
        big = [dw_gemm(dz, a, DZ_SLOTS, sv.images, b, FWD_SLOTS, S) for a, b in _BIG_PAIRS[head]]
        for li, (a, _) in zip((2, 1), _BIG_PAIRS[head][:2]):
            grads[li] = (big[2 - li], thin_dw(ones, dz, a, DZ_SLOTS, S).view(-1))
        out[head] = (grads, big[2])
    # ---- layer 0 of each head: feature columns from the GEMM, per-sample xyz / per-ray view / per-image latent columns
    W_r0, W_t0 = rgb_p[0][0], trans_p[0][0]
    dz0_r = unpack(dz, 2, DZ_SLOTS, S)
    g_ray = ops.group_colsum(dz0_r, S, N)                         # [B*R,256] per-ray sums
    g_img = ops.group_colsum(g_ray, B * R, R)                     # [B,256]
    view_t, _, vc = geom["view_seg"]()
    dW_view, _ = ops.linear_backward_weight(g_ray, [(view_t, 1, vc)], B * R, want_bias=False)           # [256,27]
    xyz = ops.points_from_depth(geom["center"], geom["ray"], geom["depth"]).view(S, 3)
    dW_xyz = thin_dw(xyz, dz, 2, DZ_SLOTS, S).t().contiguous()                                          # [256,3]
    dW_light, _ = ops.linear_backward_weight(g_img, [(ll, 1, cfg.n_latent_light)], B, want_bias=False)  # [256,48]
    grads_r, big_r0 = out["rgb"]
    grads_r[0] = (torch.cat([big_r0, dW_view, dW_xyz, dW_light], dim=1), g_img.sum(dim=0) if B > 1 else g_img.view(-1).clone())
    d_ll = ops.linear_backward_input(g_img, W_r0, B, cfg.n_latent_light, None, w_col0=256 + vc + 3) if need_lat_light else None
    dz0_t = unpack(dz, 5, DZ_SLOTS, S)
    g_timg = ops.group_colsum(dz0_t, S, per_image)                # [B,256]
    dW_lt, _ = ops.linear_backward_weight(g_timg, [(lt, 1, cfg.n_latent_trans)], B, want_bias=False)    # [256,16]
    grads_t, big_t0 = out["trans"]
    grads_t[0] = (torch.cat([big_t0, dW_lt], dim=1), g_timg.sum(dim=0) if B > 1 else g_timg.view(-1).clone())
    d_lt = ops.linear_backward_input(g_timg, W_t0, B, cfg.n_latent_trans, None, w_col0=256) if need_lat_trans else None
    return grads_r, grads_t, d_lt, d_ll
