"""Multi-GPU plumbing: one process per GPU, rays sharded by rank, table + weights replicated, and ONE
allreduce(SUM) per step over the flat fp32 gradient buffer the kernels write (replaces the reference's
``nn.DataParallel``, train.py:46-48).  Gradients are already scaled by 1/(3*B_global) on every rank, so
SUM reproduces the reference's global-mean L1 gradient and Adam stays bit-identical across ranks."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def shard_range(n: int, group=None):
    """Contiguous 1/W slice of n rays owned by this rank (ragged tail goes to the last ranks evenly)."""
    w, r = world_size(group), rank(group)
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def allreduce_flat(flat: torch.Tensor, sums: torch.Tensor, group=None):
    """Sum the flat gradient vector and the two loss sums over ranks -- ONE collective when `sums` sits directly behind
    `flat` in the same allocation (how the fused step returns them), else one each."""
    adjacent = (flat.is_contiguous() and sums.is_contiguous() and flat.dtype == sums.dtype
                and flat.untyped_storage().data_ptr() == sums.untyped_storage().data_ptr()
                and sums.storage_offset() == flat.storage_offset() + flat.numel())
    if adjacent:
        both = torch.as_strided(flat, (flat.numel() + sums.numel(),), (1,))
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)


def allreduce_module_grads(model, group=None):
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def broadcast_parameters(model, src: int = 0, group=None):
    for p in model.parameters():
        dist.broadcast(p.data, src=src, group=group)
