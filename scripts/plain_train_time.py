"""Training step of the plain model (options/nerf_lm_env.yaml) at the C3 shape -- 4096 rays x 128 samples, forward + composite +
backward -- on the tensor-core path (opt.b200.mlp = 'bf16') and on the SIMT fp32 kernels.  Never a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C  # noqa: E402
from texpose_b200.config import AttrDict, env_opt  # noqa: E402
from texpose_b200.layers.nerf import NeRF  # noqa: E402

dev = "cuda:0"
B, R, N = 16, 256, 128
g = torch.Generator().manual_seed(0)
center = (torch.randn(B, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(dev)
ray = (torch.randn(B, R, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(dev)
depth = ((torch.rand(B, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 1.2 + 0.2).to(dev)
image = torch.rand(B, R, 3, generator=g).to(dev)
for mode, steps in (("bf16", 10), ("fp32", 2)):
    opt = env_opt(device=dev)
    opt.b200 = AttrDict(mlp=mode)
    torch.manual_seed(0)
    m = NeRF(opt).to(dev)

    def step():
        for p in m.parameters():
            p.grad = None
        rgb_s, sig = m.forward_samples(opt, center, ray, depth, mode="train")
        rgb = m.composite(opt, ray, rgb_s, sig, depth)[0]
        ((rgb - image) ** 2).mean().backward()

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
    # per-entry-point device time (events around every C-ABI call, a pass of its own)
    orig, ev = _C.call, []

    def timed_call(name, *a):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); orig(name, *a); a1.record()
        ev.append((name, a0, a1))

    _C.call = timed_call
    step()
    _C.call = orig
    torch.cuda.synchronize()
    per = {}
    for name, a0, a1 in ev:
        per[name] = per.get(name, 0.0) + a0.elapsed_time(a1)
    print("   entry-point ms:", {k: round(v, 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1])}, "sum", round(sum(per.values()), 2))
    print(f"plain model, {mode}: {ms:.2f} ms per step ({B * R * N / ms / 1e3:.1f} M samples/s), C-ABI launches per step: "
          f"{ {k: v // steps for k, v in sorted(_C.launch_counts.items())} }")
