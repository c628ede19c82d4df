"""Multi-GPU plumbing for the rollout path: instances are independent, so the caller's batch is split
contiguously over the ranks (one process per GPU; augmentation copies and POMO starts stay with their
instance so the best-of reduction is local) and the only collective is ONE all-gather of the per-instance
best costs.  Training (rl.py:99-130 under upstream's DDP trainer) has one real exchange step more: the
gradients of the shards' losses are summed over the ranks, as ONE all-reduce of a flat buffer (`allreduce_gradients`).
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int):
    """Contiguous, balanced split: the first n % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    bounds, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        bounds.append((start, start + size))
        start += size
    return bounds


def shard_range(n_items: int, rank: int, world: int):
    return shard_bounds(n_items, world)[rank]


def shard_td(td, rank: int, world: int):
    """This rank's slice of an instance batch (dict of tensors with a leading batch dim)."""
    lo, hi = shard_range(next(iter(td.values())).shape[0], rank, world)
    return {k: v[lo:hi] for k, v in td.items()}


def gather_costs(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather the per-instance results of every rank back into caller order ([n_items, ...])."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    bounds = shard_bounds(n_items, world)
    sizes = [hi - lo for lo, hi in bounds]
    if len(set(sizes)) == 1:
        out = torch.empty((n_items, *local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad, *local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


def allreduce_gradients(parameters, n_local: int, n_global: int, group=None) -> None:
    """Data-parallel REINFORCE step: every rank has back-propagated the MEAN loss of its own `n_local` instances
    (`training.pomo_shared_baseline_loss` on its shard); the gradient of the mean loss over all `n_global` instances is the
    sum over the ranks of grad_r * n_local_r / n_global.  The gradients are packed into one flat buffer (one collective per
    step: launch latency, not link bandwidth, is what a ~1 MB all-reduce costs over NVSwitch), reduced, and written back in
    place.  Shards of different sizes are weighted correctly; parameters without a gradient on this rank contribute zeros."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return
    weight = float(n_local) / float(n_global)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params]) * weight
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
