"""C2-size composite forward (307 200 rays x 128): event-timed launches and achieved HBM GB/s.  Never a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import ops  # noqa: E402

DEV = "cuda:0"
R, N = 480 * 640, 128
g = torch.Generator(device=DEV).manual_seed(0)
ray = torch.randn(1, R, 3, device=DEV, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0], device=DEV)
rgb = torch.rand(1, R, N, 3, 2, device=DEV, generator=g)
den = torch.rand(1, R, N, 2, device=DEV, generator=g) * 3
depth = (torch.rand(1, R, N, 1, device=DEV, generator=g) + torch.arange(N, device=DEV)[None, None, :, None]) / N * 2.5 + 6.7
unc = torch.rand(1, R, N, 1, device=DEV, generator=g)
bytes_alg = R * N * (40 + 12) + R * 56
for _ in range(3):
    ops.CompositeSTL.apply(ray, rgb, den, depth, unc, 0.05)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.CompositeSTL.apply(ray, rgb, den, depth, unc, 0.05)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"composite_stl forward incl. output allocation: median {ts[5] * 1e3:.0f} us, best {ts[0] * 1e3:.0f} us -> "
      f"{bytes_alg / ts[5] / 1e6:.0f} GB/s algorithmic ({bytes_alg / 1e9:.2f} GB)")
