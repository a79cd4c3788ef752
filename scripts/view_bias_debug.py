"""Debugging aid: after one fused render launch the CTA scratch still holds the view-direction bias rows the kernel computed;
compare them with the tp_tc_ray_bias table of the same rays (bit-exact expected) and compare fused / unfused per-sample rgb."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C, camera, compute_box, mlp_tc, ops, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402

dev = torch.device("cuda:0")
H, W, N = 24, 40, 128
opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=str(dev))
opt.nerf.sample_stratified = False
opt.b200 = AttrDict(mlp="bf16")
torch.manual_seed(0)
g = Graph(opt, n_train_images=4).to(dev).eval()
pose, intr = synth.poses([0]).to(dev), synth.intrinsics(1).to(dev).clone()
intr[:, :2] *= 0.06
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
dr = (zn[:, :, None], zf[:, :, None])
R = 280
with torch.no_grad():
    a = g.render(opt, pose, intr=intr, ray_idx=range(80, 80 + R), depth_range=dr, mode="val")
torch.cuda.synchronize()
scratch = mlp_tc._scratch[(dev.type, dev.index)]
n_super = (R * N + 255) // 256
grid = min(n_super, 148)
vb = scratch[grid * 2 * 65536:].view(torch.float32)[: grid * 2 * 4 * 256].view(grid, 2, 4, 256)
# rows: CTA c handled super-tile c (single wave): tile 2c+t = ray 2c+t
got = vb[:, :, 0, :].reshape(grid * 2, 256)[:R]
# the table of the multi-kernel path
cfg = g.nerf._config(opt, "val")
pairs = lambda ml: [(l.weight.detach(), l.bias.detach()) for l in ml]
rgb_p, trans_p = pairs(g.nerf.mlp_rgb), pairs(g.nerf.mlp_trans)
lt = g.latent_vars_trans.weight[0][None].detach().contiguous()
ll = g.latent_vars_light.weight[0][None].detach().contiguous()
img_r, img_t = mlp_tc.image_biases(cfg, 1, lt, ll, rgb_p, trans_p)
center, ray = camera.get_center_and_ray(opt, pose, intr=intr, ray_idx=torch.arange(80, 80 + R, device=dev)[None])
W_r0 = rgb_p[0][0]
table = torch.empty(R, 256, device=dev)
_C.call("tp_tc_ray_bias", ops._p(ray), R, R, cfg.L_view, ops._p(W_r0), W_r0.stride(0), 256, ops._p(img_r), ops._p(table), ops._stream())
torch.cuda.synchronize()
d = (got - table).abs()
print("view-bias rows vs table: max abs diff", float(d.max()), "rows differing", int((d.max(dim=1).values > 0).sum()), "of", R)
if float(d.max()) > 0:
    r = int(d.max(dim=1).values.argmax())
    print(" worst row", r, "cols differing", int((d[r] > 0).sum()), "first cols", d[r].nonzero()[:8, 0].tolist(), got[r, :4].tolist(), table[r, :4].tolist())
    same_prev = (got[1:] - table[:-1]).abs().max(dim=1).values
    print(" rows equal to the PREVIOUS ray's table row:", int((same_prev == 0).sum()))
