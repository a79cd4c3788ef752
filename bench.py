#!/usr/bin/env python
"""Benchmark of the TexPose render hot path on B200 (contract: see the task's bench.py section).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): LineMOD-duck-shaped synthetic full-frame render, 480x640 rays x 128
samples, static + transient + light heads, random-init (seed 0) weights, bf16 tensor-core MLP.  One step = one
full frame through the public API (Graph.nerf_forward(mode='val')).  With N GPUs every rank renders its own
view (weak scaling over views, no data-path collective).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NS = 480, 640, 128
FLOP_PER_SAMPLE_FWD = 1_821_184          # SURVEY.md 8d (dense, unpadded)
FLOP_PER_SAMPLE_FWD_BWD = 3_220_992      # forward + backward with the trunk frozen
METRIC, UNIT = "ray-samples/sec", "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=1024, help="rays of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ---------------------------------------------------------------------------------------------- CPU reference arm

def cpu_render_sample(n_rays, threads):
    """Times the oracle (CPU restatement of the reference path: rays -> bounds -> depths -> MLP -> composite)
    on a bounded sample of the C2 workload.  Returns samples/s."""
    import torch
    from oracle import texpose_oracle as O
    from texpose_b200 import synth
    from texpose_b200.config import adapt_gan_opt
    from texpose_b200.layers.nerf_static_transient_light import NeRF

    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = NeRF(adapt_gan_opt())
    L = lambda ml: [(l.weight.detach(), l.bias.detach()) for l in ml]
    feat, rgb, trans = L(m.mlp_feat), L(m.mlp_rgb), L(m.mlp_trans)
    pose, intr = synth.poses([0]), synth.intrinsics(1)
    lo, hi = synth.padded_aabb()
    lt, ll = synth.latents(1)
    state = {}

    def step():
        with torch.no_grad():
            c, r = O.get_center_and_ray(pose, intr, H, W)
            tn, tf, v = O.aabb_ray_intersection(lo, hi, c, r)
            zn, zf = O.box_bounds_to_range(tn, tf, v, *synth.BG_RANGE)
            idx = torch.linspace(0, H * W - 1, n_rays).long()[None]
            c, r = O.gather_rays(c, idx), O.gather_rays(r, idx)
            zn, zf = zn[:, idx[0]], zf[:, idx[0]]
            rand = torch.rand(1, n_rays, NS, 1)
            state["out"] = O.render_stl(c, r, zn, zf, rand, NS, lt, ll, feat, rgb, trans)

    return step, n_rays * NS


def eager_gpu_sample(dev, n_chunks=8, chunk=2048):
    """The reference's eager aten-op path (oracle port) run ON THE GPU in its own 2048-ray chunks (opt.nerf.rand_rays,
    model/nerf_adapt_st_gan.py:669-679), cuBLAS fp32 -- the meaningful "before" on the same B200.  Reported as context only."""
    import torch
    from oracle import texpose_oracle as O
    from texpose_b200 import synth
    from texpose_b200.config import adapt_gan_opt
    from texpose_b200.layers.nerf_static_transient_light import NeRF
    torch.manual_seed(0)
    m = NeRF(adapt_gan_opt()).to(dev)
    L = lambda ml: [(l.weight.detach(), l.bias.detach()) for l in ml]
    feat, rgb, trans = L(m.mlp_feat), L(m.mlp_rgb), L(m.mlp_trans)
    pose, intr = synth.poses([0]).to(dev), synth.intrinsics(1).to(dev)
    lo, hi = [t.to(dev) for t in synth.padded_aabb()]
    lt, ll = [t.to(dev) for t in synth.latents(1)]

    def run():
        with torch.no_grad():
            for c in range(n_chunks):
                cen, ray = O.get_center_and_ray(pose, intr, H, W)          # whole frame per chunk, as the reference does
                tn, tf, v = O.aabb_ray_intersection(lo, hi, cen, ray)
                zn, zf = O.box_bounds_to_range(tn, tf, v, *synth.BG_RANGE)
                idx = torch.arange(c * chunk, (c + 1) * chunk, device=dev)[None]
                rand = torch.rand(1, chunk, NS, 1, device=dev)
                O.render_stl(O.gather_rays(cen, idx), O.gather_rays(ray, idx), zn[:, idx[0]], zf[:, idx[0]], rand, NS, lt, ll,
                             feat, rgb, trans)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(value=n_chunks * chunk * NS / (ms * 1e-3), unit=UNIT, kind="port",
                sample=f"{n_chunks} chunks x {chunk} rays x {NS} samples, eager torch ops (oracle port of the reference path) on "
                       f"this GPU, fp32 cuBLAS, full-frame ray generation per chunk as in the reference",
                ms_per_frame_extrapolated=ms * (H * W / (n_chunks * chunk)))


def run_reference(args, rank):
    import torch
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, samples = cpu_render_sample(args.cpu_rays, threads)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = samples / dt
    line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="C2 LineMOD-duck synthetic full-frame render 480x640x128, static+transient+light "
                                     "heads (bounded CPU sample)", rays_per_step=args.cpu_rays, samples_per_ray=NS),
                cpu_baseline=dict(value=val, unit=UNIT, cores=threads, kind="port",
                                  sample=f"{args.cpu_rays} evenly spaced rays x {NS} samples of the 480x640 frame per step "
                                         f"(oracle = CPU restatement of the reference path; the reference itself is a "
                                         f"Python checkout that does not travel to the GPU box)"),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                torch_threads=torch.get_num_threads())
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- clocks sampler

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                smax = max(smax, float(f[1]))
                if t0 <= t <= t1:
                    sm.append(float(f[0]))
                    for n, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                continue
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=smax or None, reasons=sorted(reasons),
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------- our arm

def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from texpose_b200 import _C, compute_box, synth
    from texpose_b200.config import AttrDict, adapt_gan_opt
    from texpose_b200.model.nerf_adapt_st_gan import Graph

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (texpose_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout while the communicator is created
        # (NCCL_DEBUG=VERSION / WARN), so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    _C.build()
    pk = peaks()

    opt = adapt_gan_opt(H=H, W=W, sample_intvs=NS, device=str(dev))
    opt.b200 = AttrDict(mlp="bf16", rng="philox")
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=8).to(dev)
    g.eval()
    # this rank's view: seed = rank (C4-style view sharding); inputs as the data loader would deliver them
    pose_h = synth.poses([rank]).pin_memory()
    intr_h = synth.intrinsics(1).pin_memory()
    lo, hi = [t.to(dev) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose_h.to(dev), intr_h.to(dev), lo, hi, H, W, *synth.BG_RANGE)
    zn_h, zf_h = zn.cpu().pin_memory(), zf.cpu().pin_memory()
    mask_h = torch.ones(1, H, W).pin_memory()
    var_dev = AttrDict(pose=pose_h.to(dev), intr=intr_h.to(dev), z_near=zn, z_far=zf, obj_mask=mask_h.to(dev),
                       idx=torch.zeros(1, dtype=torch.long, device=dev))
    samples_per_step = H * W * NS
    out_h = dict(rgb=torch.empty(1, H * W, 3).pin_memory(), depth=torch.empty(1, H * W, 1).pin_memory(),
                 opacity=torch.empty(1, H * W, 1).pin_memory(), uncert=torch.empty(1, H * W, 1).pin_memory())

    def step_resident():
        with torch.no_grad():
            return g.nerf_forward(opt, AttrDict(var_dev), mode="val")

    def step_e2e():
        with torch.no_grad():
            var = AttrDict(pose=pose_h.to(dev, non_blocking=True), intr=intr_h.to(dev, non_blocking=True),
                           z_near=zn_h.to(dev, non_blocking=True), z_far=zf_h.to(dev, non_blocking=True),
                           obj_mask=mask_h.to(dev, non_blocking=True), idx=var_dev.idx)
            ret = g.nerf_forward(opt, var, mode="val")
            for k, buf in out_h.items():
                buf.copy_(ret[k], non_blocking=True)
        return ret

    h2d = sum(t.numel() * t.element_size() for t in (pose_h, intr_h, zn_h, zf_h, mask_h))
    d2h = sum(t.numel() * t.element_size() for t in out_h.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=False):
        for _ in range(warmup):
            fn()
        barrier()
        cs = ClockSampler(local_rank) if sampler else None
        if cs:
            time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, (cs.stop(t0, t1) if cs else None)

    # --- dominant kernel (fused MLP) duration: CUDA events around the C-ABI launch on the launching stream
    kern_ms = []
    orig_call = _C.call

    def timing_call(name, *a):
        if name in ("tp_tc_nerf_stl_forward", "tp_render_fused_forward"):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            orig_call(name, *a)
            a1.record()
            kern_ms.append((a0, a1))
        else:
            orig_call(name, *a)

    _C.launch_counts.clear()
    ms, clocks = timed(step_resident, args.steps, args.warmup, sampler=True)
    launches = sum(_C.launch_counts.values()) * args.steps // (args.steps + args.warmup)
    per_step = {k: v // (args.steps + args.warmup) for k, v in _C.launch_counts.items()}
    # second pass with per-kernel events (kept out of the headline timing)
    import texpose_b200.ops as ops_mod
    import texpose_b200.mlp_tc as tc_mod
    _C.call = timing_call
    ops_mod._C.call = timing_call
    tc_mod._C.call = timing_call
    for _ in range(min(args.steps, 5)):
        step_resident()
    torch.cuda.synchronize()
    _C.call = orig_call
    kms = sorted(a.elapsed_time(b) for a, b in kern_ms)
    k_ms = sum(kms) / len(kms)
    achieved_tf = FLOP_PER_SAMPLE_FWD * samples_per_step / (k_ms * 1e-3) / 1e12

    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))

    value = world * samples_per_step / (ms * 1e-3)
    e2e_value = world * samples_per_step / (ms_e2e * 1e-3)

    train = None
    if not args.no_train:
        train = bench_train(args, g, opt, dev, world, timed)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_tc_forward_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                impl="ours",
                config=dict(workload="C2 LineMOD-duck synthetic full-frame render 480x640x128, static+transient+light heads, "
                                     "1 view per GPU per step (views sharded across ranks)",
                            rays_per_step=H * W, samples_per_ray=NS, ms_per_frame=ms, mlp="bf16 tcgen05, fp32 accumulate",
                            weights="random-init seed 0 (Xavier, as the reference)", rng="in-kernel Philox jitter",
                            l2="per-step working set ~2.2 GB (per-sample outputs + bias table) >> 126 MB L2; no flush needed"),
                roofline=dict(bound="tensor", achieved=achieved_tf, peak=pk["tf_sustained"], unit="TFLOP/s",
                              frac=achieved_tf / pk["tf_sustained"], traffic=traffic, kernel="tc::nerf_stl_forward_kernel",
                              kernel_ms=k_ms, kernel_share_of_step=k_ms / ms, peak_source=f"{pk['src']} bf16 sustained",
                              flop_per_sample=FLOP_PER_SAMPLE_FWD),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e),
                gpu_launches=launches, launches_per_step=per_step, clocks=clocks)
    if train:
        line["train_step"] = train
    if not args.no_cpu_baseline:
        try:
            line["eager_gpu_baseline"] = eager_gpu_sample(dev)
        except Exception as e:  # noqa: BLE001  (context only; never fails the bench)
            line["eager_gpu_baseline"] = dict(error=str(e)[:200])
        threads = os.cpu_count() or 1
        cstep, csamples = cpu_render_sample(args.cpu_rays, threads)
        cstep()
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            cstep()
            best = min(best, time.perf_counter() - t0)
        line["cpu_baseline"] = dict(value=csamples / best, unit=UNIT, cores=threads, kind="port",
                                    sample=f"{args.cpu_rays} evenly spaced rays x {NS} samples of the same frame, fp32, "
                                           f"best of 2 after warm-up (oracle = CPU restatement of the reference path)")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_train(args, g, opt, dev, world, timed):
    """C3: texture-learner step, 16 patches of 16x16 rays x 128 samples per GPU, fwd + bwd (+ grad allreduce)."""
    import torch
    from texpose_b200 import compute_box, parallel, synth
    from texpose_b200.config import AttrDict, adapt_gan_opt
    from texpose_b200.model.base import summarize_loss
    B, P = 16, 16
    opt_t = adapt_gan_opt(H=128, W=128, sample_intvs=NS, device=str(dev))
    opt_t.b200 = AttrDict(mlp="bf16", rng="philox")
    pose = synth.poses(list(range(B))).to(dev)
    K = torch.tensor([[572.4114, 0, 64 - 572.4114 * 0.3 / 8], [0, 573.57043, 64 + 573.57043 * 0.2 / 8], [0, 0, 1]])
    intr = K.repeat(B, 1, 1).to(dev)
    lo, hi = [t.to(dev) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, 128, 128, *synth.BG_RANGE)
    coords, _ = synth.patch_coords(B, P, seed=2)
    coords = coords.to(dev)
    idx = torch.arange(B, device=dev) % 8
    image = torch.rand(B, 3, 128, 128, device=dev)
    mask = (torch.rand(B, 128, 128, device=dev) > 0.3).float()
    params = [p for p in g.parameters() if p.requires_grad]
    # the one exchange of the path: N > 1 -> one kernel over NVLink peer windows (csrc/peer.cu); the NCCL allreduce of the same
    # bucket is timed beside it
    exchange = parallel.PeerGradExchange(params) if world > 1 else parallel.GradBucket(params)
    g.train()

    def make_step(bucket):
        def step():
            for p in params:
                p.grad = None
            ret = g.render(opt_t, pose, intr=intr, ray_idx=coords, depth_range=(zn[:, :, None], zf[:, :, None]),
                           sample_idx=idx, mode="train")
            var = AttrDict(idx=idx, image=image, obj_mask=mask, ray_idx=coords)
            var.update(ret)
            loss = g.compute_loss(opt_t, var, mode="train")      # patch gather + render / uncert / trans_reg terms, fused
            summarize_loss(opt_t, var, loss)["all"].backward()   # seeds formed on the device in the backward
            bucket.allreduce_mean()
        return step

    n_steps = max(10, args.steps)       # 2.5 ms each: enough steps for the max-over-ranks time to settle
    ms, _ = timed(make_step(exchange), n_steps, 3)
    samples = B * P * P * NS
    out = dict(workload="C3 train step fwd+bwd, 4096 rays x 128 samples per GPU, fused patch loss, bf16 tcgen05 fwd + bwd (dX chain with thin gradients, dW GEMMs), grad exchange",
               value=world * samples / (ms * 1e-3), unit="samples/s", ms_per_step=ms, grad_exchange_bytes=exchange.flat.numel() * 4,
               grad_exchange="none (1 GPU)" if world == 1 else "one-kernel rank-order mean over CUDA-IPC peer windows (NVLink)",
               tflops_per_gpu=FLOP_PER_SAMPLE_FWD_BWD * samples / (ms * 1e-3) / 1e12, flop_per_sample=FLOP_PER_SAMPLE_FWD_BWD)
    if world > 1:
        exchange.check()
        ms_nccl, _ = timed(make_step(parallel.GradBucket(params)), n_steps, 3)
        out["ms_per_step_nccl_allreduce"] = ms_nccl
        # the exchange alone, back to back (gradients already in place): device time per call, max over ranks
        for p in params:
            p.grad = torch.ones_like(p)
        out["exchange_only_us"] = {"peer": 1e3 * timed(exchange.allreduce_mean, 50, 5)[0],
                                   "nccl": 1e3 * timed(parallel.GradBucket(params).allreduce_mean, 50, 5)[0]}
        exchange.check()
        exchange.close()
    g.eval()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
