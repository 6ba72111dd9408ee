"""Multi-rank host logic (ray sharding + ONE flat-gradient allreduce + identical Adam steps) on
CPU with the gloo backend, world_size 2.  The CUDA kernels are replaced by a stand-in module whose
``l1_step`` is computed by the oracle, so this checks exactly what `parallel.py` / `trainer.py` add on top of
the kernels: shard ranges, the 1/(3*B_global) scaling, SUM-allreduce of the flat buffer and loss sums."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import field_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _OracleBackedField(torch.nn.Module):
    """Same `l1_step` contract as inf_b200.model._FusedFieldBase.l1_step, arithmetic by the CPU oracle."""

    def __init__(self, params, table, num_layers, skip):
        super().__init__()
        self.names = list(params.keys())
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(params[n].clone()) for n in self.names])
        self.table, self.num_layers, self.skip = table, num_layers, skip

    def l1_step(self, batch, global_rays=None, flat_grad=None):
        p = {n: q.detach() for n, q in zip(self.names, self.ps)}
        vids, bary, target = batch["vertex_idxs_of_hit_faces"], batch["barycentric_coords"], batch["expected_rgbs"]
        B = target.shape[0]
        _, pred, grads = O.loss_and_grads(p, self.table, vids, bary, target, num_layers=self.num_layers, skip=self.skip)
        scale = B / float(global_rays if global_rays is not None else B)     # mean over B -> mean over B_global
        flat = torch.cat([(grads[n] * scale).reshape(-1) for n in self.names])
        off = 0
        for q in self.ps:
            q.grad = flat[off:off + q.numel()].view_as(q)
            off += q.numel()
        sums = torch.stack([(pred - target).abs().sum(), ((pred - target) ** 2).sum()])
        return sums, pred, flat


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import inf_b200  # noqa: F401  (package import must work without a GPU)
        from inf_b200 import parallel, trainer
        c = load_golden("tf_k64")
        B = c["vids"].shape[0] - 1                       # 256 rays: splits evenly
        lo, hi = parallel.shard_range(B)
        model = _OracleBackedField(c["params"], c["table"], c["num_layers"], c["skip"])
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        losses = []
        for _ in range(3):
            batch = {"vertex_idxs_of_hit_faces": c["vids"][lo:hi], "barycentric_coords": c["bary"][lo:hi],
                     "expected_rgbs": c["target"][lo:hi]}
            loss, _ = trainer.train_step_async(model, opt, batch)
            losses.append(float(loss))
        out[rank] = (lo, hi, losses, [q.detach().clone() for q in model.ps])
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_step_on_the_whole_batch():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (lo0, hi0, l0, p0), (lo1, hi1, l1, p1) = out[0], out[1]
    assert (lo0, hi0, lo1, hi1) == (0, 128, 128, 256)
    assert l0 == l1                                      # the allreduced loss is the global mean on every rank
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)                         # weights stay bit-identical without a broadcast
    # single-process reference on the concatenated batch (trainer.py:71-84)
    c = load_golden("tf_k64")
    idx = [torch.arange(0, 256)] * 3
    ref, ref_losses = O.adam_train_steps(c["params"], c["table"], c["vids"], c["bary"], c["target"], idx, lr=1e-3,
                                         num_layers=6, skip=3)
    for a, n in zip(p0, ref):
        assert torch.allclose(a, ref[n], atol=2e-6, rtol=1e-4), n
    assert max(abs(x - y) for x, y in zip(l0, ref_losses)) < 1e-6


def test_shard_range_covers_ragged_counts():
    from inf_b200 import parallel
    # without a process group: world 1 owns everything
    assert parallel.shard_range(1001) == (0, 1001)
    assert parallel.world_size() == 1 and parallel.rank() == 0
