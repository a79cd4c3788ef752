"""480x640x128 frame of the plain NeRF (layers/nerf.py, nerf_lm_env.yaml dims) in rendering mode: fused tcgen05 kernel
(padded layer image) vs its fp32 kernels.  Event-timed forward_samples + composite.  Never a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200.config import AttrDict, env_opt  # noqa: E402
from texpose_b200.layers.nerf import NeRF  # noqa: E402

DEV = "cuda:0"
R, N = 480 * 640, 128
g = torch.Generator().manual_seed(4)
center = (torch.randn(1, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -8.0])).to(DEV)
ray = (torch.randn(1, R, 3, generator=g) * 0.05 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
depth = ((torch.rand(1, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 2.5 + 6.7).to(DEV)
for mlp, reps in (("bf16", 5), ("fp32", 1)):
    opt = env_opt(device=DEV, sample_intvs=N)
    opt.b200 = AttrDict(mlp=mlp, slice_rays=1 << 14)
    torch.manual_seed(0)
    m = NeRF(opt).to(DEV)
    step = R if mlp == "bf16" else 1 << 14         # the fp32 kernels keep [S,256] activations: render in slices

    def frame():
        with torch.no_grad():
            for c in range(0, R, step):
                sl = slice(c, min(c + step, R))
                rgb_s, den = m.forward_samples(opt, center[:, sl], ray[:, sl], depth[:, sl], mode="val")
                m.composite(opt, ray[:, sl], rgb_s, den, depth[:, sl])

    frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        frame()
    e1.record()
    torch.cuda.synchronize()
    print(f"plain NeRF, mlp={mlp}: {e0.elapsed_time(e1) / reps:.1f} ms per 480x640x128 frame")

# static / transient / light model, eval mode: full launch vs static-only (opt.b200.static_only)
from texpose_b200 import synth  # noqa: E402
from texpose_b200.config import adapt_gan_opt  # noqa: E402
from texpose_b200.layers.nerf_static_transient_light import NeRF as StlNeRF  # noqa: E402

lt, ll = [t.to(DEV) for t in synth.latents(1)]
for static_only in (False, True):
    opt = adapt_gan_opt(device=DEV, sample_intvs=N)
    opt.b200 = AttrDict(mlp="bf16", static_only=static_only)
    torch.manual_seed(0)
    m = StlNeRF(opt).to(DEV)

    def frame():
        with torch.no_grad():
            out = m.forward_samples(opt, center, ray, depth, lt, ll, mode="eval")
            m.composite(opt, ray, out[0], out[1], depth, out[2])

    frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        frame()
    e1.record()
    torch.cuda.synchronize()
    print(f"static/transient/light model, eval, static_only={static_only}: {e0.elapsed_time(e1) / 5:.1f} ms per 480x640x128 frame")
