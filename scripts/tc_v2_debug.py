"""v2 kernel check: outputs vs the default kernel, per-tile error map.  usage: python scripts/tc_v2_debug.py <flags>"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import mlp_tc, synth
from texpose_b200.config import adapt_gan_opt
from texpose_b200.layers import _common
from texpose_b200.layers.nerf_static_transient_light import NeRF
DEV = "cuda:0"
opt = adapt_gan_opt(device=DEV)
torch.manual_seed(0)
m = NeRF(opt).to(DEV)
B, R, N = 1, int(sys.argv[2]) if len(sys.argv) > 2 else 37, 128
g = torch.Generator().manual_seed(4)
center = (torch.randn(B, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(DEV)
ray = (torch.randn(B, R, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
depth = ((torch.rand(B, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 1.2 + 0.2).to(DEV)
lt, ll = [t.to(DEV) for t in synth.latents(B)]
cfg = m._config(opt, "val"); cfg.precision = "bf16"
geom = _common.ray_geometry(cfg, center, ray, depth)
pairs = lambda ml: [(l.weight.detach(), l.bias.detach()) for l in ml]
fp, rp, tp = pairs(m.mlp_feat), pairs(m.mlp_rgb), pairs(m.mlp_trans)
ref = mlp_tc.forward(cfg, geom, lt, ll, fp, rp, tp, flags=0)
torch.cuda.synchronize()
flags = int(sys.argv[1])
for st in (0, 1, 15):
    a = mlp_tc.forward(cfg, geom, lt, ll, fp, rp, tp, dbg_layer=st, flags=0)[3]
    b = mlp_tc.forward(cfg, geom, lt, ll, fp, rp, tp, dbg_layer=st, flags=flags)[3]
    torch.cuda.synchronize()
    err = (a - b).abs().view(R, N, 256).amax(dim=(1, 2))
    bad = (err > 1e-3).nonzero().flatten().tolist()
    print(f"stage {st}: tiles bad {len(bad)}/{R}: {bad[:40]}  max err {err.max().item():.3e}")
out = mlp_tc.forward(cfg, geom, lt, ll, fp, rp, tp, flags=flags)
torch.cuda.synchronize()
print("outputs max diff vs default kernel:", [(a - b).abs().max().item() for a, b in zip(out, ref)])
