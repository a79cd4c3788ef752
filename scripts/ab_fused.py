"""Same-box A/B of one C2 frame (480x640x128, bf16): fused render launch against the multi-kernel path it replaces.
Alternating rounds, CUDA events, frames resident.  usage: python scripts/ab_fused.py [rounds] [frames]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C, compute_box, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H, W, N = 480, 640, 128
dev = "cuda:0"


def make(fused):
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=dev)
    opt.b200 = AttrDict(mlp="bf16", rng="philox", fused_render=fused)
    return opt


torch.manual_seed(0)
g = Graph(make(True), n_train_images=8).to(dev).eval()
pose, intr = synth.poses([0]).to(dev), synth.intrinsics(1).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
var = AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=torch.ones(1, H, W, device=dev), idx=torch.zeros(1, dtype=torch.long, device=dev))


def run(opt, n):
    with torch.no_grad():
        for _ in range(3):
            g.nerf_forward(opt, AttrDict(var), mode="val")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            g.nerf_forward(opt, AttrDict(var), mode="val")
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {True: [], False: []}
for r in range(rounds):
    for fused in (True, False):
        res[fused].append(run(make(fused), frames))
print("fused   ms/frame:", " ".join(f"{x:.2f}" for x in res[True]))
print("unfused ms/frame:", " ".join(f"{x:.2f}" for x in res[False]))
