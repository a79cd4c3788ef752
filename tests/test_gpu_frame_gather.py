"""One frame rendered by several ranks (parallel.FrameGather): every rank's fused launch stores its row
block of the per-ray outputs straight into the root's frame buffers (CUDA-IPC window; NVLink stores when the ranks sit on
different GPUs), one barrier kernel per frame.  The gathered frame must equal the single-GPU frame BIT FOR BIT (same kernel,
same per-ray arithmetic), for consecutive frames (the two window buffers alternate)."""
import os
import socket

import pytest
import torch

from texpose_b200 import _C

pytestmark = pytest.mark.gpu
H, W, N = 32, 64, 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from texpose_b200 import compute_box, parallel, synth
    from texpose_b200.config import AttrDict, adapt_gan_opt
    from texpose_b200.model.nerf_adapt_st_gan import Graph
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fg = None
    try:
        opt = adapt_gan_opt(H=H, W=W, sample_intvs=N, device=str(dev))
        opt.nerf.sample_stratified = False
        opt.b200 = AttrDict(mlp="bf16")
        torch.manual_seed(0)
        g = Graph(opt, n_train_images=2).to(dev).eval()
        lo, hi = [t.to(dev) for t in synth.padded_aabb()]
        fg = parallel.FrameGather(opt, device=dev, timeout_ms=30000)
        ok = True
        for frame, seed in enumerate((0, 3, 5)):          # three frames: both window buffers, one reused
            pose, intr = synth.poses([seed]).to(dev), synth.intrinsics(1).to(dev).clone()
            intr[:, :2] *= 0.1
            zn, zf = compute_box.box_range(pose, intr, lo, hi, H, W, *synth.BG_RANGE)
            dr = (zn[:, :, None], zf[:, :, None])
            with torch.no_grad():
                got, local = fg.render(g, opt, pose, intr, dr, local_keys=("alpha_static",))
                b, e = fg.rows
                assert local["alpha_static"].shape == (1, e - b, N)
                torch.cuda.synchronize(dev)
                fg.check()
                if rank == 0:
                    want = g.render(opt, pose, intr=intr, ray_idx=range(0, H * W), depth_range=dr, mode="val")
                    for k in parallel.PER_RAY_KEYS:
                        ok = ok and torch.equal(got[k], want[k])
                    ok = ok and torch.equal(local["alpha_static"], want["alpha_static"][:, b:e])
                else:
                    assert got is None
        ret[rank] = bool(ok)
    finally:
        if fg is not None:
            fg.close()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_frame_gather_equals_single_gpu_frame(world):
    import time
    import torch.multiprocessing as mp
    _C.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=False)
    deadline = time.time() + 300
    done = False
    while not done and time.time() < deadline:
        done = ctx.join(timeout=5)
    if not done:
        for p in ctx.processes:
            if p.is_alive():
                p.terminate()
        pytest.fail("frame-gather workers did not finish")
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_frame_layout_offsets():
    from texpose_b200 import parallel
    layout, total = parallel.frame_layout(480 * 640)
    off = 0
    for k in parallel.PER_RAY_KEYS:
        o, c = layout[k]
        assert o == off and o % 64 == 0
        off += (480 * 640 * c + 63) // 64 * 64
    assert total == off and total * 4 < 20 * 2 ** 20      # 14 floats per ray: 17 MB per frame buffer
