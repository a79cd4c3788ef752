"""Cycle counters of the fused forward kernel's MMA warp (library built with TEXPOSE_NVCC_EXTRA=-DTP_FWD_PROF): per stage, how
long the tensor pipe's issuer waited for the tile's epilogue warps and for weight chunks.  Debugging aid, never a benchmark.
usage: TEXPOSE_B200_LIB=build/libtexpose_prof.so python scripts/fwd_prof.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import _C, compute_box, mlp_tc, synth  # noqa: E402
from texpose_b200.config import AttrDict, adapt_gan_opt  # noqa: E402
from texpose_b200.model.nerf_adapt_st_gan import Graph  # noqa: E402

dev = torch.device("cuda:0")
H, W, N = 480, 640, 128
torch.manual_seed(0)
opt0 = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=str(dev))
opt0.b200 = AttrDict(mlp="bf16", rng="philox")
g = Graph(opt0, n_train_images=8).to(dev).eval()
pose, intr = synth.poses([0]).to(dev), synth.intrinsics(1).to(dev)
lo, hi = [t.to(dev) for t in synth.padded_aabb()]
zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
var = AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=torch.ones(1, H, W, device=dev), idx=torch.zeros(1, dtype=torch.long, device=dev))
lib = _C.load()
off = lib.tp_tc_prof_offset()
for fused in (True, False, True, False):
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=str(dev))
    opt.b200 = AttrDict(mlp="bf16", rng="philox", fused_render=fused)
    with torch.no_grad():
        for _ in range(3):
            g.nerf_forward(opt, AttrDict(var), mode="val")
    torch.cuda.synchronize()
    scratch = mlp_tc._scratch[(dev.type, dev.index)]
    c = scratch[off:].view(torch.int64)[: 148 * 2 * 17 * 4].view(148, 2, 17, 4).double()
    n_st = c[:, 0, 0, 3].mean().item()
    total = c[:, 0, 0, 2].mean().item()
    print(f"{'fused' if fused else 'per-sample'} kernel: {total / n_st:9.0f} cycles per super-tile ({n_st:.0f} super-tiles per CTA)")
    rdy = c[:, :, :, 0].mean(dim=0) / n_st       # [tile, stage]
    full = c[:, 0, :, 1].mean(dim=0) / n_st
    print("  stage:            " + " ".join(f"{L:5d}" for L in range(17)))
    print("  wait epilogue T0: " + " ".join(f"{v:5.0f}" for v in rdy[0].tolist()) + f"   sum {rdy[0].sum():7.0f}")
    print("  wait epilogue T1: " + " ".join(f"{v:5.0f}" for v in rdy[1].tolist()) + f"   sum {rdy[1].sum():7.0f}")
    print("  wait weights:     " + " ".join(f"{v:5.0f}" for v in full.tolist()) + f"   sum {full.sum():7.0f}")
