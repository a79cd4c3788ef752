"""Data-parallel texture-learner step at the reference yaml's step size (8 patches x 256 rays x 64 samples PER GPU), one process per
GPU: eager (render, fused loss, backward, gradient allreduce, Adam) against the same step captured as one CUDA graph per rank
(texpose_b200.train_graph.GraphedStep with the NCCL allreduce of parallel.GradBucket inside the capture).  Checks that all ranks hold
the same weights afterwards.  torchrun --nproc-per-node N scripts/graph_dp_step.py.  Never a benchmark."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import compute_box, parallel, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.base import summarize_loss  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402
from texpose_b200.train_graph import GraphedStep  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, P, N = 8, 16, 64
opt = adapt_gan_opt(H=128, W=128, sample_intvs=N, device=str(dev))
opt.batch_size = B
opt.b200 = AttrDict(mlp="bf16", rng="torch")
pose = synth.poses(list(range(rank * B, rank * B + B))).to(dev)            # every rank its own views and patches
K = torch.tensor([[572.4114, 0, 64 - 572.4114 * 0.3 / 8], [0, 573.57043, 64 + 573.57043 * 0.2 / 8], [0, 0, 1]])
intr = K.repeat(B, 1, 1).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, 128, 128, *synth.BG_RANGE)
coords = synth.patch_coords(B, P, seed=2 + rank)[0].to(dev)
idx = torch.arange(B, device=dev)
gen = torch.Generator().manual_seed(10 + rank)
image = torch.rand(B, 3, 128, 128, generator=gen).to(dev)
mask = (torch.rand(B, 128, 128, generator=gen) > 0.3).float().to(dev)


def build(capturable):
    torch.manual_seed(0)                                                   # same weights on every rank
    g = Graph(opt, n_train_images=B).to(dev).train()
    op = torch.optim.Adam([dict(params=g.nerf.parameters(), lr=1.e-3)], capturable=capturable)
    op.add_param_group(dict(params=g.latent_vars_light.parameters(), lr=1.e-3))
    op.add_param_group(dict(params=g.latent_vars_trans.parameters(), lr=1.e-3))
    bucket = parallel.GradBucket([p for p in g.parameters() if p.requires_grad])

    def fwd_bwd():
        ret = g.render(opt, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]), sample_idx=idx, mode="train")
        var = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
        var.update(ret)
        total = summarize_loss(opt, var, g.compute_loss(opt, var, mode="train"))["all"]
        total.backward()
        bucket.allreduce_mean()
        return total

    return g, op, fwd_bwd


def timed(fn, steps=40, warm=5):
    for _ in range(warm):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def spread(g):
    """max over parameters of |mine - rank 0's|"""
    worst = 0.0
    for p in g.parameters():
        ref = p.detach().clone()
        if world > 1:
            dist.broadcast(ref, 0)
        worst = max(worst, float((ref - p.detach()).abs().max()))
    return worst


g_e, op_e, fb_e = build(False)


def eager():
    op_e.zero_grad(set_to_none=True)
    fb_e()
    op_e.step()


ms_e = timed(eager)
g_g, op_g, fb_g = build(True)
step = GraphedStep(fb_g, op_g, static=dict(coords=coords, image=image, mask=mask), warmup=3)
ms_g = timed(lambda: step(coords=coords, image=image, mask=mask))
s_e, s_g = spread(g_e), spread(g_g)
if rank == 0:
    n = world * B * P * P * N
    print(f"{world} GPU(s), {B} x {P * P} rays x {N} samples per GPU and step, gradient allreduce inside the step:")
    print(f"  eager  {ms_e:.2f} ms per step ({n / ms_e / 1e3:.1f} M samples/s), weights differ across ranks by at most {s_e:.1e}")
    print(f"  graph  {ms_g:.2f} ms per step ({n / ms_g / 1e3:.1f} M samples/s), weights differ across ranks by at most {s_g:.1e}")
# a process group whose NCCL communicator was captured into a live CUDA graph does not tear down cleanly (the first run of this script
# hung in destroy_process_group): synchronise, then leave without running the teardown
sys.stdout.flush()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
os._exit(0)
