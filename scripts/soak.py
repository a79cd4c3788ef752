"""Soak run: the same C3 training step / C2 frame / parity-mode strip / plain-model step repeated many times must reproduce their
first result bit for bit (every kernel here is deterministic by construction; a rare race in a hand-written barrier protocol
would show up as a differing repeat).  usage: python scripts/soak.py [train_reps] [frame_reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import compute_box, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt, env_opt  # noqa: E402
from texpose_b200.layers.nerf import NeRF as PlainNeRF  # noqa: E402
from texpose_b200.model.base import summarize_loss  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402

n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_frame = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
H, W, NS = 480, 640, 128
opt = adapt_gan_opt(H=H, W=W, sample_intvs=NS, device=str(dev))
opt.nerf.sample_stratified = False
opt.b200 = AttrDict(mlp="bf16")
torch.manual_seed(0)
g = Graph(opt, n_train_images=8).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]

# ---- C3 training step, fixed inputs (midpoint depths: no RNG)
B, P = 16, 16
opt_t = adapt_gan_opt(H=128, W=128, sample_intvs=NS, device=str(dev))
opt_t.nerf.sample_stratified = False
opt_t.b200 = AttrDict(mlp="bf16")
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


def train_step():
    for p in params:
        p.grad = None
    ret = g.render(opt_t, pose_t, intr=intr_t, ray_idx=coords, depth_range=(znt[:, :, None], zft[:, :, None]), sample_idx=idx, mode="train")
    v = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
    v.update(ret)
    summarize_loss(opt_t, v, g.compute_loss(opt_t, v, mode="train"))["all"].backward()
    return torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])


first = train_step().clone()
bad = 0
for i in range(n_train):
    if not torch.equal(train_step(), first):
        bad += 1
print(f"C3 training step: {n_train} repeats, {bad} differ from the first (finite: {bool(torch.isfinite(first).all())})")

# ---- C2 frame (fused render launch) and a parity-mode strip (split kernel)
g.eval()
pose, intr = synth.poses([0]).to(dev), synth.intrinsics(1).to(dev)
zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
var = AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=torch.ones(1, H, W, device=dev), idx=torch.zeros(1, dtype=torch.long, device=dev))


def frame(o, rays=None):
    with torch.no_grad():
        if rays is None:
            out = g.nerf_forward(o, AttrDict(var), mode="val")
        else:
            out = g.render(o, pose, intr=intr, ray_idx=rays, depth_range=(zn[:, :, None], zf[:, :, None]), mode="val")
    return torch.cat([out[k].reshape(-1) for k in ("rgb", "depth", "opacity", "uncert")])


f0 = frame(opt).clone()
bad = sum(0 if torch.equal(frame(opt), f0) else 1 for _ in range(n_frame))
print(f"C2 frame (render launch): {n_frame} repeats, {bad} differ")
o32 = AttrDict(opt)
o32.b200 = AttrDict(mlp="fp32")
strip = torch.arange(200 * W, 200 * W + 16384, device=dev)[None]
s0 = frame(o32, strip).clone()
bad = sum(0 if torch.equal(frame(o32, strip), s0) else 1 for _ in range(n_frame))
print(f"parity-mode strip (split kernel, 16384 rays): {n_frame} repeats, {bad} differ")

# ---- plain-model training step
oe = env_opt(device=str(dev))
oe.b200 = AttrDict(mlp="bf16")
torch.manual_seed(0)
pm = PlainNeRF(oe).to(dev)
gg = torch.Generator().manual_seed(0)
c_ = (torch.randn(16, 256, 3, generator=gg) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(dev)
r_ = (torch.randn(16, 256, 3, generator=gg) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(dev)
d_ = ((torch.rand(16, 256, NS, 1, generator=gg) + torch.arange(NS)[None, None, :, None]) / NS * 1.2 + 0.2).to(dev)


def plain_step():
    for p in pm.parameters():
        p.grad = None
    rgb_s, sig = pm.forward_samples(oe, c_, r_, d_, mode="train")
    (pm.composite(oe, r_, rgb_s, sig, d_)[0].square().mean() + 0.01 * sig.mean()).backward()
    return torch.cat([p.grad.reshape(-1) for p in pm.parameters()])


p0 = plain_step().clone()
bad = sum(0 if torch.equal(plain_step(), p0) else 1 for _ in range(n_train // 2))
print(f"plain-model training step: {n_train // 2} repeats, {bad} differ")
