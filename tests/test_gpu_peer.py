"""Gradient exchange over peer windows (csrc/peer.cu): the one-kernel publish / wait / rank-order reduce.

The protocol is exercised three ways: (1) several "ranks" of ONE process on one device, each on its own stream (the
kernels wait for each other exactly as ranks on different GPUs do, minus NVLink); (2) a rank that never shows up -> bounded
wait, status word, output poisoned with NaN; (3) two PROCESSES exchanging through real CUDA IPC handles (on two GPUs when the box has
them, else both on cuda:0) against the torch.distributed allreduce of the same gradients."""
import os
import socket

import pytest
import torch

from texpose_b200 import _C, parallel

pytestmark = pytest.mark.gpu


def _ranks_in_one_process(world, n, epochs, grid):
    dev = torch.device("cuda:0")
    wins = [parallel.PeerWindow(n, dev) for _ in range(world)]
    ptrs = [w.ptr for w in wins]
    outs = [torch.full(((n + 3) // 4 * 4,), -7.0, device=dev) for _ in range(world)]
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    g = torch.Generator(device=dev).manual_seed(n + world)
    try:
        for e in range(1, epochs + 1):
            grads = [torch.randn(n, device=dev, generator=g) * (10.0 ** (r % 3)) for r in range(world)]
            for r in range(world):
                wins[r].buffers[e & 1].copy_(grads[r])
            torch.cuda.synchronize()
            order = list(range(world)) if e % 2 else list(reversed(range(world)))      # launch order must not matter
            for r in order:
                parallel.peer_allreduce_mean(ptrs, r, n, e, outs[r], grid_ctas=grid, timeout_ms=5000, stream=streams[r])
            torch.cuda.synchronize()
            want = grads[0].cpu()
            for r in range(1, world):
                want = want + grads[r].cpu()
            want = (want / world).to(dev)        # IEEE division as on the CPU (torch's CUDA `/ scalar` multiplies by 1/world)
            for r in range(world):
                assert wins[r].status() == 0
                assert torch.equal(outs[r][:n], want), (world, n, e, r)      # bit-exact: same order, IEEE add / divide
    finally:
        for w in wins:
            w.close()


@pytest.mark.parametrize("world,n,grid", [(2, 421385, 0), (4, 1000, 4), (8, 421385, 8), (3, 7, 1), (1, 513, 2)])
def test_rank_order_mean_bit_exact(world, n, grid):
    # grids are kept small when many "ranks" share one GPU so that every kernel's CTA 0 is resident while the others wait
    _ranks_in_one_process(world, n, epochs=4, grid=grid if grid else 16)


def test_no_cta_waits_for_its_own_grid():
    """A grid far larger than one resident wave (1 rank: nothing to wait for; 2 ranks on one device, launched back to back):
    no CTA depends on CTA 0 of its own grid being resident, so the exchange completes whatever the grid size."""
    _ranks_in_one_process(1, 1 << 20, epochs=2, grid=4096)
    _ranks_in_one_process(2, 1 << 18, epochs=2, grid=64)


def test_missing_peer_times_out_and_reports():
    dev = torch.device("cuda:0")
    n = 4096
    wins = [parallel.PeerWindow(n, dev) for _ in range(2)]
    try:
        out = torch.full((n,), 3.0, device=dev)
        wins[0].buffers[1].fill_(1.0)
        parallel.peer_allreduce_mean([w.ptr for w in wins], 0, n, 1, out, grid_ctas=4, timeout_ms=50)
        torch.cuda.synchronize()
        assert wins[0].status() == 1 and wins[1].status() == 0
        assert torch.isnan(out).all()        # poisoned: stale / partial gradients can never reach an optimizer step unnoticed
    finally:
        for w in wins:
            w.close()


def test_argument_errors():
    dev = torch.device("cuda:0")
    w = parallel.PeerWindow(64, dev)
    try:
        out = torch.zeros(64, device=dev)
        with pytest.raises(RuntimeError):
            parallel.peer_allreduce_mean([w.ptr], 0, 64, 0, out)              # epochs start at 1
        with pytest.raises(RuntimeError):
            parallel.peer_allreduce_mean([w.ptr], 1, 64, 1, out)              # rank outside the world
        with pytest.raises(RuntimeError):
            parallel.peer_allreduce_mean([w.ptr], 0, 64, 1, out[1:])          # too small / misaligned
        with pytest.raises(RuntimeError):
            parallel.PeerWindow(8, "cpu")
    finally:
        w.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = None
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(37, 19).to(dev)
        emb = torch.nn.Embedding(5, 3).to(dev)
        params = list(lin.parameters()) + list(emb.parameters())
        ex = parallel.PeerGradExchange(params, timeout_ms=20000)
        ok = True
        for step in range(3):
            g = torch.Generator(device=dev).manual_seed(100 * step + rank)
            grads = [torch.randn(p.shape, device=dev, generator=g) for p in params]
            if rank == 1 and step == 1:
                grads[2] = None                                     # an embedding no sample touched: counts as zero
            for p, gr in zip(params, grads):
                p.grad = gr
            flat = torch.cat([(gr if gr is not None else torch.zeros_like(p)).reshape(-1) for p, gr in zip(params, grads)])
            parts = [torch.empty_like(flat).cpu() for _ in range(world)]
            dist.all_gather(parts, flat.cpu())
            want = parts[0].clone()
            for q in parts[1:]:
                want = want + q
            want = (want / world).to(dev)
            ex.allreduce_mean()
            torch.cuda.synchronize()
            ex.check()
            got = torch.cat([p.grad.reshape(-1) for p in params])
            ok = ok and torch.equal(got, want)
        ret[rank] = bool(ok)
    finally:
        if ex is not None:
            ex.close()
        dist.destroy_process_group()


def test_two_processes_through_cuda_ipc():
    import torch.multiprocessing as mp
    _C.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.spawn(_ipc_worker, args=(2, port, ret), nprocs=2, join=False)
    import time
    deadline = time.time() + 240
    done = False
    while not done and time.time() < deadline:
        done = ctx.join(timeout=5)          # True once every worker has exited; raises if one failed
    if not done:
        for p in ctx.processes:
            if p.is_alive():
                p.terminate()
        pytest.fail("IPC exchange workers did not finish")
    assert ret.get(0) and ret.get(1)
