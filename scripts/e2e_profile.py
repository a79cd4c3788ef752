"""Where the end-to-end frame (host buffers in, host buffers out) spends its time beside the resident frame.  Never a benchmark."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import compute_box, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model.nerf_adapt_st_gan import Graph
from torch.profiler import ProfilerActivity, profile
dev = torch.device("cuda:0")
H, W, NS = 480, 640, 128
opt = adapt_gan_opt(H=H, W=W, sample_intvs=NS, device=str(dev))
opt.b200 = AttrDict(mlp="bf16", rng="philox")
torch.manual_seed(0)
g = Graph(opt, n_train_images=8).to(dev).eval()
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
pose_h, intr_h = synth.poses([0]).pin_memory(), synth.intrinsics(1).pin_memory()
zn, zf = compute_box.box_range(pose_h.to(dev), intr_h.to(dev), lo, hi, H, W, *synth.BG_RANGE)
zn_h, zf_h, mask_h = zn.cpu().pin_memory(), zf.cpu().pin_memory(), torch.ones(1, H, W).pin_memory()
idx0 = torch.zeros(1, dtype=torch.long, device=dev)
out_h = dict(rgb=torch.empty(1, H * W, 3).pin_memory(), depth=torch.empty(1, H * W, 1).pin_memory(),
             opacity=torch.empty(1, H * W, 1).pin_memory(), uncert=torch.empty(1, H * W, 1).pin_memory())

def step_e2e():
    with torch.no_grad():
        var = AttrDict(pose=pose_h.to(dev, non_blocking=True), intr=intr_h.to(dev, non_blocking=True),
                       z_near=zn_h.to(dev, non_blocking=True), z_far=zf_h.to(dev, non_blocking=True),
                       obj_mask=mask_h.to(dev, non_blocking=True), idx=idx0)
        ret = g.nerf_forward(opt, var, mode="val")
        for k, buf in out_h.items():
            buf.copy_(ret[k], non_blocking=True)

for _ in range(3):
    step_e2e()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step_e2e()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev = t0
for e in ev:
    print(f"{e.time_range.start - t0:10.1f} us  dur {e.time_range.end - e.time_range.start:10.1f}  gap {e.time_range.start - prev:8.1f}  {e.name[:80]}")
    prev = max(prev, e.time_range.end)
