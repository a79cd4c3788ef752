"""The pre-training engine's Graph (model/nerf_pretrain.py) at the shapes of options/nerf_lm_env.yaml -- training step of
2048 rays x 64 samples (forward + compute_loss + backward), validation frame 480 x 640 x 64 through render_by_slices -- for the
unmodified reference (eager torch, from baseline/_ref) and for texpose_b200.model.nerf_pretrain.Graph on the same GPU.
Never a benchmark; numbers go to DESIGN.md."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402
from texpose_b200 import compute_box, synth  # noqa: E402
from texpose_b200.config import AttrDict  # noqa: E402
from texpose_b200.model import nerf_pretrain as BP  # noqa: E402

dev = "cuda:0"
H, W, B = 480, 640, 2
ns = ref_import.load()
mod = importlib.import_module("model.nerf_pretrain")
opt = ref_import.load_yaml_opt("nerf_lm_env", H, W, device=dev)
opt.loss_weight.update(render=0, mask=-1, depth=-1)
opt.data.erode_mask_loss = False
pose = synth.poses([0, 1]).to(dev)
intr = synth.intrinsics(B).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
gen = torch.Generator().manual_seed(3)
var0 = dict(idx=torch.arange(B), pose=pose, pose_init=pose, intr=intr, z_near=zn, z_far=zf,
            image=torch.rand(B, 3, H, W, generator=gen).to(dev), obj_mask=(zf < 29).view(B, H, W).float(),
            depth_gt=(7.5 + torch.rand(B, H, W, generator=gen)).to(dev))


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(name, graph, o, steps_train, steps_frame):
    def train():
        for p in graph.parameters():
            p.grad = None
        var = graph.forward(o, AttrDict(var0), mode="train")
        loss = graph.compute_loss(o, var, mode="train")
        sum(10 ** float(o.loss_weight[k]) * loss[k] for k in loss).backward()

    def frame():
        with torch.no_grad():
            graph.render_by_slices(o, pose[:1], intr=intr[:1], depth_range=(zn[:1, :, None], zf[:1, :, None]),
                                   object_mask=var0["obj_mask"][:1], mode="val")

    t, f = timed(train, steps_train), timed(frame, steps_frame, warm=1)
    n_t, n_f = o.nerf.rand_rays * o.nerf.sample_intvs, H * W * o.nerf.sample_intvs
    print(f"{name}: training step {t:.2f} ms ({n_t / t / 1e3:.1f} M samples/s), validation frame {f:.1f} ms ({n_f / f / 1e3:.1f} M samples/s)")
    return t, f


torch.manual_seed(0)
g_ref = mod.Graph(opt).to(dev)
ref = run("reference, eager torch", g_ref, opt, 10, 2)
for mlp in ("bf16", "fp32"):
    o = AttrDict(opt)
    o.b200 = AttrDict(mlp=mlp)
    g = BP.Graph(o).to(dev)
    g.load_state_dict(g_ref.state_dict())
    ours = run(f"texpose_b200 ({mlp})", g, o, 20, 5)
    print(f"   x{ref[0] / ours[0]:.1f} training step, x{ref[1] / ours[1]:.1f} validation frame")
