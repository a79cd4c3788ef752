"""Host-issue time vs device time of one C2 render step and one C3 training step (is a step launch-bound?).

  python scripts/step_timeline.py [steps]

Prints, per workload: ms of host time to ISSUE a step (no synchronisation inside the loop), ms of device time per step
(CUDA events around the same loop) and the sum of the launches the C-ABI saw.  Never a benchmark number.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C, compute_box, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402
from texpose_b200.model.base import summarize_loss  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")


def measure(name, fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _C.launch_counts.clear()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name}: host issue {1e3 * (t1 - t0) / steps:.3f} ms/step, device {e0.elapsed_time(e1) / steps:.3f} ms/step, "
          f"{sum(_C.launch_counts.values()) // steps} C-ABI launches/step", flush=True)


# ---- C2 render
H, W, NS = 480, 640, 128
opt = adapt_gan_opt(H=H, W=W, sample_intvs=NS, device=str(dev))
opt.b200 = AttrDict(mlp="bf16", rng="philox")
torch.manual_seed(0)
g = Graph(opt, n_train_images=8).to(dev)
g.eval()
pose, intr = synth.poses([0]).to(dev), synth.intrinsics(1).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
var = AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=torch.ones(1, H, W, device=dev),
               idx=torch.zeros(1, dtype=torch.long, device=dev))


def render():
    with torch.no_grad():
        g.nerf_forward(opt, AttrDict(var), mode="val")


measure("C2 render", render, steps)
if os.environ.get("TP_TORCH_PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            render()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    t_first = ev[0].time_range.start
    print("device activity timeline of the profiled steps (start us, dur us, gap-before us, name):")
    prev_end = t_first
    for e in ev:
        print(f"{e.time_range.start - t_first:10.1f} {e.time_range.end - e.time_range.start:10.1f} {e.time_range.start - prev_end:8.1f}  {e.name[:90]}")
        prev_end = max(prev_end, e.time_range.end)

# ---- C3 training step
B, P = 16, 16
opt_t = adapt_gan_opt(H=128, W=128, sample_intvs=NS, device=str(dev))
opt_t.b200 = AttrDict(mlp="bf16", rng="philox")
pose_t = synth.poses(list(range(B))).to(dev)
K = torch.tensor([[572.4114, 0, 64 - 572.4114 * 0.3 / 8], [0, 573.57043, 64 + 573.57043 * 0.2 / 8], [0, 0, 1]])
intr_t = K.repeat(B, 1, 1).to(dev)
znt, zft = compute_box.box_range(pose_t, intr_t, lo, hi, 128, 128, *synth.BG_RANGE)
coords = synth.patch_coords(B, P, seed=2)[0].to(dev)
idx = torch.arange(B, device=dev) % 8
image = torch.rand(B, 3, 128, 128, device=dev)
mask = (torch.rand(B, 128, 128, device=dev) > 0.3).float()
params = [p for p in g.parameters() if p.requires_grad]
g.train()


def train():
    for p in params:
        p.grad = None
    ret = g.render(opt_t, pose_t, intr=intr_t, ray_idx=coords, depth_range=(znt[:, :, None], zft[:, :, None]),
                   sample_idx=idx, mode="train")
    v = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
    v.update(ret)
    summarize_loss(opt_t, v, g.compute_loss(opt_t, v, mode="train"))["all"].backward()


measure("C3 train", train, steps)
if os.environ.get("TP_TORCH_PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            train()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=32, max_name_column_width=60))
