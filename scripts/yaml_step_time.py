"""Texture-learner training step at the shape of options/nerf_lm_adapt_gan.yaml itself -- batch_size 8 patches of 16 x 16 rays x 64
samples = 131 072 samples, a quarter of BASELINE's C3 -- with the engine's Adam step: device time per step, host time to issue a
step, and the unmodified reference's Graph.render + the same three ray-wise loss terms on the same GPU.  Never a benchmark."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402
from texpose_b200 import _C, compute_box, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.base import summarize_loss  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402

dev = "cuda:0"
B, P, N = 8, 16, 64
opt = adapt_gan_opt(H=128, W=128, sample_intvs=N, device=dev)
opt.batch_size = B
opt.b200 = AttrDict(mlp="bf16", rng="torch")
pose = synth.poses(list(range(B))).to(dev)
K = torch.tensor([[572.4114, 0, 64 - 572.4114 * 0.3 / 8], [0, 573.57043, 64 + 573.57043 * 0.2 / 8], [0, 0, 1]])
intr = K.repeat(B, 1, 1).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, 128, 128, *synth.BG_RANGE)
coords = synth.patch_coords(B, P, seed=2)[0].to(dev)
idx = torch.arange(B, device=dev)
image = torch.rand(B, 3, 128, 128, device=dev)
mask = (torch.rand(B, 128, 128, device=dev) > 0.3).float()


def adam(graph):
    o = torch.optim.Adam([dict(params=graph.nerf.parameters(), lr=1.e-3)])      # model/nerf_adapt_st_gan.py:62-69
    o.add_param_group(dict(params=graph.latent_vars_light.parameters(), lr=1.e-3))
    o.add_param_group(dict(params=graph.latent_vars_trans.parameters(), lr=1.e-3))
    return o


def measure(name, step, steps=30):
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    dev_ms, host_ms = e0.elapsed_time(e1) / steps, 1e3 * (t1 - t0) / steps
    print(f"{name}: {dev_ms:.2f} ms per step on the device ({B * P * P * N / dev_ms / 1e3:.1f} M samples/s), host issues a step in {host_ms:.2f} ms")
    return dev_ms


torch.manual_seed(0)
g = Graph(opt, n_train_images=B).to(dev).train()
optim = adam(g)


def ours():
    optim.zero_grad()
    ret = g.render(opt, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]), sample_idx=idx, mode="train")
    var = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
    var.update(ret)
    loss = g.compute_loss(opt, var, mode="train")
    summarize_loss(opt, var, loss)["all"].backward()
    optim.step()


_C.launch_counts.clear()
t_ours = measure("texpose_b200 (bf16), yaml step 8 x 256 rays x 64", ours)
print("   C-ABI launches per step:", sum(_C.launch_counts.values()) // 35)

# ---- the same step captured once in a CUDA graph (every C-ABI launch goes to the caller's stream, nothing synchronises or allocates
# outside torch's caching allocator, so the whole step -- render, fused loss, backward, Adam(capturable=True), weight re-pack -- replays
# as one graph launch; the coords / image / pose buffers are static and would be refilled in place between replays)
try:
    torch.manual_seed(0)
    g2 = Graph(opt, n_train_images=B).to(dev).train()
    optim2 = torch.optim.Adam([dict(params=g2.nerf.parameters(), lr=1.e-3)], capturable=True)
    optim2.add_param_group(dict(params=g2.latent_vars_light.parameters(), lr=1.e-3))
    optim2.add_param_group(dict(params=g2.latent_vars_trans.parameters(), lr=1.e-3))

    def body():
        ret = g2.render(opt, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]), sample_idx=idx, mode="train")
        var = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
        var.update(ret)
        loss = g2.compute_loss(opt, var, mode="train")
        total = summarize_loss(opt, var, loss)["all"]
        total.backward()
        optim2.step()
        return total

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            optim2.zero_grad(set_to_none=True)
            body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    optim2.zero_grad(set_to_none=True)
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        static_loss = body()
    before = [p.detach().clone() for p in g2.nerf.mlp_rgb.parameters()]
    t_graph = measure("texpose_b200 (bf16), the same step replayed as one CUDA graph", cg.replay)
    moved = max(float((a - b.detach()).abs().max()) for a, b in zip(before, g2.nerf.mlp_rgb.parameters()))
    print(f"   loss after the replays {float(static_loss):.4f}, head weights moved by up to {moved:.2e} (the optimizer ran inside the graph)")
except Exception as e:  # noqa: BLE001
    t_graph = None
    print("   CUDA-graph capture of the step failed:", repr(e)[:300])

if ref_import.available():
    ns = ref_import.load()
    ropt = ref_import.load_yaml_opt("nerf_lm_adapt_gan", 128, 128, device=dev)
    ropt.batch_size = B
    g_ref = ref_import.build_graph(ns, ropt, n_images=B, seed=0).to(dev).train()
    for p in g_ref.nerf.mlp_feat.parameters():
        p.requires_grad_(False)
    roptim = adam(g_ref)
    import torch.nn.functional as F

    def ref():
        roptim.zero_grad()
        ret = g_ref.render(ropt, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]), sample_idx=idx, mode="train")
        img = F.grid_sample(image, coords, mode="bilinear", align_corners=True).view(B, 3, -1).permute(0, 2, 1)
        m = F.grid_sample(mask[:, None], coords, mode="nearest").view(B, 1, -1).permute(0, 2, 1)
        render = (m * (img - ret.rgb) ** 2 / ret.uncert ** 2).sum() / (m.sum() + 1e-5)      # :747-763, the three ray-wise terms
        uncert = 5 + torch.log(ret.uncert ** 2).mean() / 2
        reg = ret.density[..., -1].mean()
        (render + uncert + 0.01 * reg).backward()
        roptim.step()

    t_ref = measure("reference, eager torch, same step", ref, steps=10)
    print(f"   x{t_ref / t_ours:.1f}" + (f" eager, x{t_ref / t_graph:.1f} as a graph" if t_graph else ""))
