"""Cycle counters of the CTA-pair forward kernel (control warps + one epilogue warp per tile).  Debugging aid, never a benchmark.
Build with  TEXPOSE_NVCC_EXTRA=-DTP_PAIR_PROF python -c "from texpose_b200 import _C; _C.build(force=True)"  first.
usage: python scripts/pair_prof.py <flags> [rays]"""
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
flags = int(sys.argv[1])
B, R, N = 1, int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 2 * 2 * 16, 128
g = torch.Generator().manual_seed(4)
center = (torch.randn(B, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(DEV)
ray = (torch.randn(B, R, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
depth = ((torch.rand(B, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 1.2 + 0.2).to(DEV)
lt, ll = [t.to(DEV) for t in synth.latents(B)]
cfg = m._config(opt, "val"); cfg.precision = "bf16"
geom = _common.ray_geometry(cfg, center, ray, depth)
pairs = lambda ml: [(l.weight.detach(), l.bias.detach()) for l in ml]
fp, rp, tp = pairs(m.mlp_feat), pairs(m.mlp_rgb), pairs(m.mlp_trans)
for _ in range(2):
    out = mlp_tc.forward(cfg, geom, lt, ll, fp, rp, tp, dbg_layer=100, flags=flags)
torch.cuda.synchronize()
c = out[3].view(-1)[:148 * 16 * 2].view(torch.int64).view(148, 16).cpu()
lead, peer = c[0::2].double(), c[1::2].double()
stages = lead[:, 7] * 17
f = lambda t, i: (t[:, i] / stages).mean().item()
print("flags", flags, "per stage (both tiles), leader ctrl warp: total %.0f | wait go %.0f | in elect: MMA issue %.0f commits %.0f | wait weights %.0f, own A %.0f | issue window %.0f"
      % tuple(f(lead, i) for i in range(7)))
print("   epilogue warp 0 (both tiles): leader wait acc %.0f work %.0f | peer wait acc %.0f work %.0f" % (f(lead, 8), f(lead, 9), f(peer, 8), f(peer, 9)))

tr = out[3].view(-1).view(torch.int64)[148 * 16:148 * 16 + 2 * 17 * 2 * 8].view(2, 17, 2, 8).cpu()
lead = tr[0]
base = int(lead[1, 0, 0])
print("leader CTA, iteration 3, cycles relative to stage 1 / tile 0 first MMA.  columns: first MMA | last issue | epilogue woke | drain done | own A ready seen | peer A ready seen")
for L in range(1, 12):
    for t in range(2):
        e = lead[L, t] - base
        print("L%2d t%d  mma0 %6d  last %6d (+%5d)  epi woke %6d (+%5d)  drained %6d (+%5d)  own ready seen %6d (+%5d)  peer ready %6d (+%5d)  go %6d (+%4d)" % (
            L, t, e[0], e[1], e[1] - e[0], e[2], e[2] - e[1], e[3], e[3] - e[2], e[6], e[6] - e[3], e[4], e[4] - e[6], e[5], e[5] - e[4]))
