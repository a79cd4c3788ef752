"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (the data path itself needs no collective for
rendering; training has one grad allreduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from texpose_b200 import parallel


def test_shard_partitions_cover_everything_once():
    for n in (0, 1, 7, 64, 307200, 19424):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                b, e = parallel.shard_range(n, r, world)
                assert 0 <= b <= e <= n
                got += list(range(b, e)) if n < 1000 else [(b, e)]
            if n < 1000:
                assert got == list(range(n))
            else:
                assert got[0][0] == 0 and got[-1][1] == n and all(a[1] == b[0] for a, b in zip(got, got[1:]))
    assert [parallel.shard_views(64, r, 8) for r in range(8)][3] == list(range(24, 32))
    b, e = parallel.shard_rays(480 * 640, 1, 4, align=640)
    assert b % 640 == 0 and e % 640 == 0 and (b, e) == (120 * 640, 240 * 640)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        emb = torch.nn.Embedding(4, 2)
        frozen = torch.nn.Linear(2, 2)
        for p in frozen.parameters():
            p.requires_grad = False
        params = list(lin.parameters()) + list(emb.parameters()) + list(frozen.parameters())
        bucket = parallel.GradBucket(params)
        assert bucket.flat.numel() == 15 + 3 + 8
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.full_like(lin.bias, 10.0 * (rank + 1))
        # emb.weight.grad stays None on rank 1 (no sample touched it)
        if rank == 0:
            emb.weight.grad = torch.arange(8.0).view(4, 2)
        bucket.allreduce_mean()
        ok = torch.allclose(lin.weight.grad, torch.full_like(lin.weight, 1.5)) and \
            torch.allclose(lin.bias.grad, torch.full_like(lin.bias, 15.0)) and \
            torch.allclose(emb.weight.grad, torch.arange(8.0).view(4, 2) / 2)
        # uneven ray shards gathered back in order
        sizes = [3, 5]
        b, e = (0, 3) if rank == 0 else (3, 8)
        local = torch.arange(b, e, dtype=torch.float32)[:, None].repeat(1, 2)
        full = parallel.gather_ray_outputs(local, sizes)
        ok = ok and torch.equal(full[:, 0], torch.arange(8.0))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_grad_bucket_allreduce_and_gather_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
