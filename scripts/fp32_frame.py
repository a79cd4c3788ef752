"""Frame time of the fp32 (<= 1e-4 parity) mode on the C2 frame -- split-bf16 tensor-core kernel vs the SIMT kernels -- beside the
bf16 modes, for the record.  Never a benchmark."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texpose_b200 import compute_box, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model.nerf_adapt_st_gan import Graph
dev = torch.device("cuda:0")
H, W, NS = 480, 640, 128
modes = (("fp32", "auto"),) if os.environ.get("TP_SPLIT_ONLY") else (("fp32", "auto"), ("fp32", "simt"), ("bf16", "auto"), (None, "auto"))
for mode, engine in modes:
    opt = adapt_gan_opt(H=H, W=W, sample_intvs=NS, device=str(dev))
    opt.b200 = AttrDict(rng="philox", fp32_engine=engine) if mode is None else AttrDict(mlp=mode, rng="philox", fp32_engine=engine)
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=8).to(dev).eval()
    pose, intr = synth.poses([0]).to(dev), synth.intrinsics(1).to(dev)
    lo, hi = [t.to(dev) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
    var = AttrDict(pose=pose, intr=intr, z_near=zn, z_far=zf, obj_mask=torch.ones(1, H, W, device=dev), idx=torch.zeros(1, dtype=torch.long, device=dev))
    with torch.no_grad():
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = g.nerf_forward(opt, AttrDict(var), mode="val")
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{mode} (fp32 engine {engine}): {dt * 1e3:.1f} ms/frame ({H * W * NS / dt / 1e6:.1f} M samples/s), rgb mean {out.rgb.mean().item():.4f}")
