"""One C3-shaped training step (fwd + bwd) for ncu launch lists.  Never a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import compute_box, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402
from texpose_b200.model.base import summarize_loss  # noqa: E402

dev = torch.device("cuda:0")
B, P, NS = 16, 16, 128
opt = adapt_gan_opt(H=128, W=128, sample_intvs=NS, device=str(dev))
opt.b200 = AttrDict(mlp="bf16", rng="philox")
torch.manual_seed(0)
g = Graph(opt, n_train_images=8).to(dev)
pose = synth.poses(list(range(B))).to(dev)
K = torch.tensor([[572.4114, 0, 64 - 572.4114 * 0.3 / 8], [0, 573.57043, 64 + 573.57043 * 0.2 / 8], [0, 0, 1]])
intr = K.repeat(B, 1, 1).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, 128, 128, *synth.BG_RANGE)
coords = synth.patch_coords(B, P, seed=2)[0].to(dev)
idx = torch.arange(B, device=dev) % 8
image = torch.rand(B, 3, 128, 128, device=dev)
mask = (torch.rand(B, 128, 128, device=dev) > 0.3).float()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for p in g.parameters():
        p.grad = None
    ret = g.render(opt, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]), sample_idx=idx, mode="train")
    var = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
    var.update(ret)
    loss = summarize_loss(opt, var, g.compute_loss(opt, var, mode="train"))["all"]
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
