"""Short single-GPU run of the fused forward for ncu (8 super-tiles per SM).  Never a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.layers.nerf_static_transient_light import NeRF  # noqa: E402

DEV = "cuda:0"
opt = adapt_gan_opt(device=DEV)
opt.b200 = AttrDict(mlp="bf16")
torch.manual_seed(0)
m = NeRF(opt).to(DEV)
R, N = 148 * 16, 128                    # 2368 rays x 128 = 1184 super-tiles = 8 per SM
g = torch.Generator().manual_seed(4)
center = (torch.randn(1, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -8.0])).to(DEV)
ray = (torch.randn(1, R, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
depth = ((torch.rand(1, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 2.5 + 6.7).to(DEV)
lt, ll = [t.to(DEV) for t in synth.latents(1)]
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        out = m.forward_samples(opt, center, ray, depth, lt, ll, mode="val")
        comp = m.composite(opt, ray, *out[:2], depth, out[2])
torch.cuda.synchronize()
print("ok", float(comp[0].mean()))
