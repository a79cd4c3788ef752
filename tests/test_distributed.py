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


# ---------------------------------------------------------------------------------------------------------------------
# the whole data-parallel training step on two gloo ranks (host logic; the kernel wrappers are the oracle's restatements)

def _dp_problem():
    from texpose_b200 import synth
    from texpose_b200.config import AttrDict, adapt_gan_opt
    B, P, HW, N = 4, 6, 64, 16
    opt = adapt_gan_opt(H=HW, W=HW, sample_intvs=N)
    opt.nerf.sample_stratified = False
    pose = synth.poses(list(range(B)))
    K = torch.tensor([[286.2, 0, 32 - 286.2 * 0.3 / 8], [0, 286.8, 32 + 286.8 * 0.2 / 8], [0, 0, 1]])
    intr = K.repeat(B, 1, 1)
    gen = torch.Generator().manual_seed(4)
    data = AttrDict(pose=pose, intr=intr, z_near=torch.full((B, HW * HW), 7.0), z_far=torch.full((B, HW * HW), 9.0),
                    coords=synth.patch_coords(B, P, seed=3)[0], image=torch.rand(B, 3, HW, HW, generator=gen),
                    mask=(torch.rand(B, HW, HW, generator=gen) > 0.3).float(), idx=torch.arange(B))
    return opt, data


def _dp_step(opt, data, views, exchange=False):
    """One texture-learner step on `views` of the batch: Graph.render(mode='train') -> compute_loss -> summarize_loss -> backward
    (-> the gradient exchange); returns the flat gradient of the trainable parameters (heads + latent tables)."""
    from texpose_b200.config import AttrDict
    from texpose_b200.model.base import summarize_loss
    from texpose_b200.model.nerf_adapt_st_gan import Graph
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=4)
    params = [p for p in g.parameters() if p.requires_grad]
    v = slice(views[0], views[-1] + 1)
    ret = g.render(opt, data.pose[v], intr=data.intr[v], ray_idx=data.coords[v],
                   depth_range=(data.z_near[v][:, :, None], data.z_far[v][:, :, None]), sample_idx=data.idx[v], mode="train")
    var = AttrDict(idx=data.idx[v], image=data.image[v], obj_mask=data.mask[v], ray_idx=data.coords[v])
    var.update(ret)
    summarize_loss(opt, var, g.compute_loss(opt, var, mode="train"))["all"].backward()
    if exchange:
        parallel.GradBucket(params).allreduce_mean()
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])


def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracle_swap
        oracle_swap.install()                                   # this process only
        torch.set_num_threads(2)
        opt, data = _dp_problem()
        views = parallel.shard_views(len(data.pose), rank, world)
        ret[rank] = _dp_step(opt, data, views, exchange=True).numpy()
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_step_equals_the_mean_of_the_shard_gradients_gloo_world2(monkeypatch):
    """the exchanged gradients of a 2-rank step == the mean of the per-shard gradients computed one after the other in
    one process, and the two ranks end up with the same gradient (so their weights stay identical)."""
    from tests import oracle_swap
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    got0, got1 = torch.from_numpy(ret[0]), torch.from_numpy(ret[1])
    assert torch.equal(got0, got1)
    oracle_swap.install(monkeypatch.setattr)
    opt, data = _dp_problem()
    shards = [_dp_step(opt, data, parallel.shard_views(4, r, 2)) for r in range(2)]
    want = (shards[0] + shards[1]) / 2
    assert want.abs().max() > 1e-4 and (shards[0] - shards[1]).abs().max() > 1e-5      # the shards really differ
    assert (got0 - want).abs().max() <= 1e-6 * max(1.0, float(want.abs().max()))
