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


class PeerExchange:
    """The multi-GPU training step's gradient exchange over NVLink peer memory (include/inf_field.h: InfComm).

    Every rank creates one exchange region in the C library, the ranks swap the 64-byte CUDA IPC handles with one
    ``all_gather_object`` (setup only -- no collective runs per step) and map each other's regions.  The fused training
    step (``FusedFieldAdam.fused_l1_step``) then sums the per-rank gradients inside its reduction kernel, in rank order,
    so every rank applies bit-identical updates without NCCL and without a separate optimiser launch.
    One node only (IPC); every rank must take the same sequence of steps."""

    _registry = {}

    def __init__(self, cfg, device, group=None):
        import ctypes as C
        from . import _lib, ops
        self.lib = lib = _lib.load()
        self.world, self.rank = world_size(group), rank(group)
        if not 2 <= self.world <= _lib.MAX_PEERS:
            raise ValueError(f"PeerExchange needs 2..{_lib.MAX_PEERS} ranks on one node, got {self.world}")
        self.device = torch.device(device)
        nbytes = int(lib.inf_comm_region_bytes(C.byref(ops._desc(cfg, 0, 0))))
        region, handle = C.c_void_p(), C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            _lib.check(lib.inf_comm_create(nbytes, C.byref(region), handle), "inf_comm_create")
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            regions = (C.c_void_p * _lib.MAX_PEERS)()
            self._mapped = []
            for q, h in enumerate(handles):
                if q == self.rank:
                    regions[q] = region.value
                else:
                    m = C.c_void_p()
                    _lib.check(lib.inf_comm_open(C.create_string_buffer(h, _lib.IPC_HANDLE_BYTES), C.byref(m)), "inf_comm_open")
                    regions[q] = m.value
                    self._mapped.append(m)
            dist.barrier(group=group)           # every rank has mapped every region before the first step
        self._region = region
        self.struct = _lib.InfComm(world=self.world, rank=self.rank, regions=regions, region_bytes=nbytes)
        self.id = len(PeerExchange._registry) + 1
        PeerExchange._registry[self.id] = self

    @staticmethod
    def lookup(comm_id: int):
        return PeerExchange._registry.get(comm_id)

    def close(self):
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for m in self._mapped:
                self.lib.inf_comm_close(m)
            self._mapped = []
            if self._region is not None:
                self.lib.inf_comm_free(self._region)
                self._region = None
        PeerExchange._registry.pop(self.id, None)
