"""GPU check of the tensor-core backward kernels against torch (bf16-emulating).  timeout 300 python scripts/tc_bwd_debug.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import mlp_tc, mlp_tc_bwd as bw, synth  # noqa: E402
from texpose_b200.config import adapt_gan_opt  # noqa: E402
from texpose_b200.layers import _common  # noqa: E402
from texpose_b200.layers.nerf_static_transient_light import NeRF  # noqa: E402

DEV = "cuda:0"
bf = lambda t: t.bfloat16().float()
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
cfg = m._config(opt, "train")
cfg.precision = "bf16"
geom = _common.ray_geometry(cfg, center, ray, depth)
pairs = lambda ml: [(l.weight.detach(), l.bias.detach()) for l in ml]
feat_p, rgb_p, trans_p = pairs(m.mlp_feat), pairs(m.mlp_rgb), pairs(m.mlp_trans)
S = B * R * N
rgb, den, unc, images = mlp_tc.forward(cfg, geom, lt, ll, feat_p, rgb_p, trans_p, save=True)
torch.cuda.synchronize()
H = [bw.unpack(images, s, 7, S) for s in range(7)]
print("saved activations: min/max", [f"{h.min().item():.2f}/{h.max().item():.2f}" for h in H])
dz_rgb = (torch.randn(S, 3, generator=g) * 0.1).to(DEV)
dz_trans = (torch.randn(S, 5, generator=g) * 0.1).to(DEV)
packed = bw.pack_bwd(None, rgb_p, trans_p)
dz = bw.backward_chain(dz_rgb, dz_trans, S, packed, images)
torch.cuda.synchronize()
print("chain kernel ran")
for head, layers, dz3, hs, base in (("rgb", rgb_p, dz_rgb, (H[3], H[2], H[1]), 0), ("trans", trans_p, dz_trans, (H[6], H[5], H[4]), 3)):
    d = bf(dz3)
    for i, (W, h) in enumerate(zip((layers[3][0], layers[2][0], layers[1][0]), hs)):
        d = bf((d @ bf(W)) * (h > 0))
        got = bw.unpack(dz, base + i, 6, S)
        print(f"{head} dz{2 - i}: max|ref| {d.abs().max().item():.4f}  max err {(got - d).abs().max().item():.3e}")
for flags in (0, 1):
    try:
        for name, a, b in (("rgb W2", 0, 2), ("trans W1", 4, 4), ("rgb W0 feat", 2, 0)):
            got = bw.dw_gemm(dz, 6, images, 7, [(a, b), (a, b)], S, flags=flags)[1]
            torch.cuda.synchronize()
            ref = bw.unpack(dz, a, 6, S).t() @ H[b]
            print(f"flags={flags} dW {name}: max|ref| {ref.abs().max().item():.4f}  max err {(got - ref).abs().max().item():.3e}")
    except Exception as e:  # noqa: BLE001
        print("flags", flags, "failed:", str(e)[:200])
        break
t = bw.thin_dw(dz_trans, images, 6, 7, S)
print("thin dW trans3: err", (t - dz_trans.t() @ H[6]).abs().max().item(), "ref", (dz_trans.t() @ H[6]).abs().max().item())
o = bw.thin_dw(torch.ones(S, 1, device=DEV), dz, 1, 6, S)
print("colsum via thin: err", (o.view(-1) - bw.unpack(dz, 1, 6, S).sum(0)).abs().max().item())
rs = bw.image_ray_sums(dz, 2, 6, S, N)
print("ray sums: err", (rs - bw.unpack(dz, 2, 6, S).view(B * R, N, 256).sum(1)).abs().max().item())
print("thin colsum: err", (bw.thin_colsum(dz_trans, S) - dz_trans.sum(0)).abs().max().item())
