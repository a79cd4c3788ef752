"""C3-shaped training step (16 x 256 rays x 128 samples, forward + composite + backward of both heads and the latents) of a
static / transient / light model whose heads are NOT the yaml's 3 x 256 -- i.e. on the staged kernels -- beside the yaml's own
architecture on the lock-step kernel + fused backward and beside the SIMT fp32 kernels.  Never a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.layers.nerf_static_transient_light import NeRF  # noqa: E402

dev = "cuda:0"
B, R, N = 16, 256, 128
g = torch.Generator().manual_seed(0)
center = (torch.randn(B, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(dev)
ray = (torch.randn(B, R, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(dev)
depth = ((torch.rand(B, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 1.2 + 0.2).to(dev)
image = torch.rand(B, R, 3, generator=g).to(dev)
lt, ll = [t.to(dev).requires_grad_(True) for t in synth.latents(B)]
cases = [("yaml architecture, lock-step kernel + fused backward", {}, "bf16", 10),
         ("heads 2 x 256 (staged kernels)", dict(layers_rgb=[None, 256, 256, 3], layers_trans=[None, 256, 256, 5]), "bf16", 10),
         ("heads 2 x 256, fp32 SIMT kernels", dict(layers_rgb=[None, 256, 256, 3], layers_trans=[None, 256, 256, 5]), "fp32", 2)]
for name, arch, mode, steps in cases:
    opt = adapt_gan_opt(device=dev)
    for k, v in arch.items():
        opt.arch[k] = v
    opt.b200 = AttrDict(mlp=mode)
    torch.manual_seed(0)
    m = NeRF(opt).to(dev)

    def step():
        for p in list(m.parameters()) + [lt, ll]:
            p.grad = None
        rgb_s, den, unc = m.forward_samples(opt, center, ray, depth, lt, ll, mode="train")
        comp = m.composite(opt, ray, rgb_s, den, depth, unc)
        (((comp[0] - image) ** 2 / comp[8] ** 2).mean() + torch.log(comp[8] ** 2).mean() + 0.01 * den[..., 1].mean()).backward()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _C.launch_counts.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name}: {ms:.2f} ms per step ({B * R * N / ms / 1e3:.1f} M samples/s); launches {sum(_C.launch_counts.values()) // steps}")
