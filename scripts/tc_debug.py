"""GPU debugging aid for csrc/mlp_tc.cu: compares every stage of the fused tcgen05 forward with a torch chain that
emulates bf16 operands / fp32 accumulation.  Run on the B200 box:  timeout 300 python scripts/tc_debug.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import mlp_tc, ops, synth  # noqa: E402
from texpose_b200.config import adapt_gan_opt  # noqa: E402
from texpose_b200.layers import _common  # noqa: E402
from texpose_b200.layers.nerf_static_transient_light import NeRF  # noqa: E402

DEV = "cuda:0"
bf = lambda t: t.bfloat16().float()


def reference_chain(m, center, ray, depth, lt, ll, N):
    B, R = center.shape[:2]
    x = center[:, :, None] + ray[:, :, None] * depth                      # [B,R,N,3]
    freq = 2 ** torch.arange(10, dtype=torch.float32, device=DEV) * math.pi
    s = x[..., None] * freq
    enc = torch.cat([x, torch.stack([s.sin(), s.cos()], -2).flatten(-3)], -1).view(-1, 63)
    enc_b = bf(enc)
    acts = {}
    W = lambda l: bf(l.weight.detach())
    h = enc_b
    f = m.mlp_feat
    stage = 0
    for li in range(7):
        inp = torch.cat([h, enc_b], -1) if li == 4 else h
        h = bf(torch.relu(inp @ W(f[li]).T + f[li].bias))
        acts[li] = h
    z7 = h @ W(f[7]).T + f[7].bias
    sigma_s = torch.nn.functional.softplus(z7[:, 0])
    feat = bf(torch.relu(z7[:, 1:]))
    acts[8] = feat
    unit = torch.nn.functional.normalize(ray, dim=-1)
    sv = unit[..., None] * (2 ** torch.arange(4, dtype=torch.float32, device=DEV) * math.pi)
    venc = torch.cat([unit, torch.stack([sv.sin(), sv.cos()], -2).flatten(-3)], -1)      # [B,R,27]
    r0 = m.mlp_rgb[0]
    Wr = r0.weight.detach()
    raybias = r0.bias + venc @ Wr[:, 256:283].T + (ll @ Wr[:, 286:].T)[:, None]              # [B,R,256] fp32
    raybias = raybias[:, :, None].expand(B, R, N, 256).reshape(-1, 256)
    h = bf(torch.relu(feat @ bf(Wr[:, :256]).T + enc_b[:, :3] @ bf(Wr[:, 283:286]).T + raybias))
    acts[9] = h
    for i, st in ((1, 10), (2, 11)):
        h = bf(torch.relu(h @ W(m.mlp_rgb[i]).T + m.mlp_rgb[i].bias))
        acts[st] = h
    rgb_s = torch.sigmoid(h @ W(m.mlp_rgb[3]).T + m.mlp_rgb[3].bias)
    t0 = m.mlp_trans[0]
    Wt = t0.weight.detach()
    imgb = (t0.bias + lt @ Wt[:, 256:].T)[:, None, None].expand(B, R, N, 256).reshape(-1, 256)
    h = bf(torch.relu(feat @ bf(Wt[:, :256]).T + imgb))
    acts[13] = h
    for i, st in ((1, 14), (2, 15)):
        h = bf(torch.relu(h @ W(m.mlp_trans[i]).T + m.mlp_trans[i].bias))
        acts[st] = h
    o = h @ W(m.mlp_trans[3]).T + m.mlp_trans[3].bias
    out = dict(rgb_s=rgb_s, rgb_t=torch.sigmoid(o[:, :3]), sigma_s=sigma_s,
               sigma_t=torch.nn.functional.softplus(o[:, 3]), unc=torch.nn.functional.softplus(o[:, 4]))
    return acts, out


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    opt = adapt_gan_opt(device=DEV)
    torch.manual_seed(0)
    m = NeRF(opt).to(DEV)
    B, R, N = 2, 37, 64
    g = torch.Generator().manual_seed(4)
    center = (torch.randn(B, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(DEV)
    ray = (torch.randn(B, R, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
    depth = ((torch.rand(B, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 1.2 + 0.2).to(DEV)
    lt, ll = [t.to(DEV) for t in synth.latents(B)]
    cfg = m._config(opt, "val")
    cfg.precision = "bf16"
    geom = _common.ray_geometry(cfg, center, ray, depth)
    pairs = lambda ml: [(l.weight.detach(), l.bias.detach()) for l in ml]
    feat_p, rgb_p, trans_p = pairs(m.mlp_feat), pairs(m.mlp_rgb), pairs(m.mlp_trans)
    acts, ref = reference_chain(m, center, ray, depth, lt, ll, N)
    stages = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 13, 14, 15]
    flag_list = [int(a) for a in sys.argv[1:]] or [0]
    for flags in flag_list:
        print(f"==== flags={flags}")
        for st in stages:
            rgb, den, unc, dbg = mlp_tc.forward(cfg, geom, lt, ll, feat_p, rgb_p, trans_p, dbg_layer=st, flags=flags)
            torch.cuda.synchronize()
            err = (dbg - acts[st]).abs().max().item()
            print(f"stage {st:2d}: max|act diff| = {err:.4e}   ref max {acts[st].abs().max().item():.3f}  "
                  f"nan={int(torch.isnan(dbg).sum())}", flush=True)
        e = dict(rgb_s=(rgb[:, :, 0] - ref["rgb_s"]).abs().max().item(),
                 rgb_t=(rgb[:, :, 1] - ref["rgb_t"]).abs().max().item(),
                 sigma_s=(den[:, 0] - ref["sigma_s"]).abs().max().item(),
                 sigma_t=(den[:, 1] - ref["sigma_t"]).abs().max().item(), unc=(unc - ref["unc"]).abs().max().item())
        print("outputs vs bf16-emulated chain:", {k: f"{v:.3e}" for k, v in e.items()})
        cfg32 = m._config(opt, "val")
        cfg32.precision = "fp32"
        from texpose_b200.layers._mlp import mlp_forward_fp32
        r32, d32, u32 = mlp_forward_fp32(cfg32, geom["enc"](), geom["view_seg"](), lt, ll, geom["S"], geom["per_image"],
                                         feat_p, rgb_p, trans_p, None)
        print("outputs vs fp32 path: rgb %.3e density %.3e uncert %.3e" % (
            (rgb - r32).abs().max().item(), (den - d32).abs().max().item(), (unc - u32).abs().max().item()))


if __name__ == "__main__":
    main()
