"""Multi-GPU plumbing for the render path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

The reference is single-GPU (options.py:112).  The path shards naturally:
  * rendering: views (or contiguous row blocks of one frame) are partitioned across ranks, no exchange;
    an optional all_gather collects the 56 B/ray outputs;
  * training: data-parallel over patches; ONE exchange per step -- a sum-allreduce of a flat fp32 buffer with
    the head + embedding gradients (~1.7 MB), then a divide by the world size.
Samples of one ray are never split (scan dependence).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n` items for `rank`; sizes differ by at most one, order preserved."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    b, e = shard_range(n_views, rank, world)
    return list(range(b, e))


def shard_rays(n_rays: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """Row-block partition of one frame's rays; `align` keeps shard boundaries on multiples (e.g. image rows)."""
    units = (n_rays + align - 1) // align
    b, e = shard_range(units, rank, world)
    return min(b * align, n_rays), min(e * align, n_rays)


class GradBucket:
    """Flat fp32 buffer over the trainable parameters (heads + latent embeddings): one allreduce per step."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)

    def pack(self):
        """One launch: concatenate the gradients into the flat buffer (missing gradients count as zero)."""
        base = self.flat.untyped_storage().data_ptr()
        parts = []
        for p in self.params:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if g.untyped_storage().data_ptr() == base:      # a view left by the previous unpack, accumulated into: no aliasing
                g = g.clone()
            parts.append(g.reshape(-1))
        if parts:
            torch.cat(parts, out=self.flat)
        return self.flat

    def unpack(self):
        """No copy back: every .grad becomes a view into the reduced flat buffer."""
        off = 0
        for p, n in zip(self.params, self.sizes):
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def allreduce_mean(self, group=None):
        """Mean over the ranks of every gradient (per-shard losses are means): one concatenation, one allreduce."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        self.pack()
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:                                       # gloo (CPU tests) has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)
        self.unpack()


def gather_ray_outputs(local: torch.Tensor, sizes: Sequence[int], group=None) -> torch.Tensor:
    """all_gather of per-ray outputs [R_local, C] from uneven contiguous shards -> [sum(sizes), C] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    m = max(sizes)
    pad = torch.zeros(m, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)], dim=0)
