"""Training-step surface of the reference's ``Trainer`` (trainer.py:71-84, 158-187) on the fused path.

``train_step`` keeps ``_train_step``'s contract -- ``(loss.item(), pred_rgbs)`` after forward, loss,
zero_grad, backward and ``optim.step()`` -- and uses the fused forward+backward kernel when the loss is
the config's L1 (config.py:117-118); any other loss callable goes through autograd into the same
backward kernels.  Epoch loops, TensorBoard, checkpoints stay the reference's host code.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops, parallel


def batch_rays(batch) -> int:
    """Rays in a batch: its ``ray_idxs`` when the selection is folded into the kernels, else its target rows."""
    r = batch.get("ray_idxs")
    return int(r.shape[0]) if r is not None else int(batch["expected_rgbs"].shape[0])


def _launched(batch):
    """Tell a prefetching loader that this batch's kernels are enqueued (HostRayLoader issues its next copies now)."""
    cb = batch.get("on_launched")
    if cb is not None:
        cb()


def get_loss_fn(loss_type: str) -> Callable:
    """config.py:113-122."""
    if loss_type == "L2":
        return F.mse_loss
    if loss_type == "L1":
        return F.l1_loss
    if loss_type == "cauchy":
        return lambda rgb_pred, rgb_gt: ((20 / 255) * (20 / 255) * torch.log(
            1 + (rgb_pred - rgb_gt) ** 2 / ((20 / 255) * (20 / 255)))).mean()
    raise RuntimeError(f"Unknown loss function: {loss_type}. Please use either 'L1', 'L2' or 'cauchy'")


def train_step_async(model, optim, batch, loss_fn=F.l1_loss, group=None, flat_grad=None):
    """One optimisation step without a host sync; returns (loss tensor (0-dim, device), pred)."""
    world = parallel.world_size(group)
    B = batch_rays(batch)
    if loss_fn is F.l1_loss:
        optim.zero_grad(set_to_none=True)
        sums, pred, flat = model.l1_step(batch, global_rays=B * world, flat_grad=flat_grad)
        if world > 1:
            parallel.allreduce_flat(flat, sums, group)
        loss = sums[0] / (3.0 * B * world)
    else:
        pred = model(batch)
        loss = loss_fn(pred, batch["expected_rgbs"])
        optim.zero_grad(set_to_none=True)
        (loss / world).backward()
        if world > 1:
            parallel.allreduce_module_grads(model, group)
    optim.step()
    _launched(batch)
    return loss, pred


def train_step(model, optim, batch, loss_fn=F.l1_loss, group=None) -> Tuple[float, torch.Tensor]:
    """trainer.py:71-84: returns (loss.item(), pred_rgbs)."""
    loss, pred = train_step_async(model, optim, batch, loss_fn, group)
    return loss.item(), pred


@torch.no_grad()
def evaluate(model, data_loader, loss_fn=F.l1_loss):
    """trainer.py:164-187 without the TensorBoard writes: (val_loss, epoch PSNR)."""
    model.eval()
    acc_loss = acc_l2 = 0.0
    total = 0
    for batch in data_loader:
        pred = model(batch)
        gt = batch["expected_rgbs"] if batch.get("ray_idxs") is None else batch["expected_rgbs"][batch["ray_idxs"].long()]
        n = gt.size()[0]
        acc_l2 += F.mse_loss(pred, gt, reduction="sum").item()
        acc_loss += loss_fn(pred, gt).item() * n
        total += n
    return acc_loss / total, epoch_psnr(acc_l2 / total)


def epoch_psnr(x: float) -> float:
    """evaluation_metrics.py:25-26."""
    return -10.0 * math.log10(x)


class FusedFieldAdam:
    """``torch.optim.Adam(model.parameters(), lr)`` (config.py:108; default betas/eps, no weight decay) whose update runs
    INSIDE the gradient-reduction kernel of the fused training step, together with the refresh of the fp16 operand
    images: a single-GPU optimisation step is three kernel launches (forward/backward, wgrad, reduce+Adam+pack).

    The model's parameters are re-pointed to views of ONE flat fp32 buffer in state_dict order (so ``state_dict()``,
    checkpoints and ``parameters()`` keep working); do not move the model to another device afterwards.
    ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's layout (per-parameter ``step``, ``exp_avg``,
    ``exp_avg_sq``), so optimiser checkpoints are interchangeable (trainer.py:224-227,309-311)."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8):
        self.model = model
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        params = model._ordered_params()
        layout = ops.param_layout(model.cfg)
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedFieldAdam needs the model on a CUDA device (no CPU fallback)")
        self.flat = torch.empty(ops.num_params(model.cfg), dtype=torch.float32, device=dev)
        with torch.no_grad():
            for (off, r, c), p in zip(layout, params):
                view = self.flat[off:off + r * c].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # t, on the device: a graph replay advances it
        self.layout = layout
        self.packed = ops.pack_weights(model.cfg, [p.detach() for p in params], None)     # zero padding written once

    # ---- the two ways a step is taken ---------------------------------------------------------------------------
    def fused_l1_step(self, batch, global_rays=None):
        """Single rank: forward + L1 + backward + Adam + operand refresh; returns (loss_sums, pred, flat_grad)."""
        m = self.model
        table, vids, bary, dense = m._inputs(batch)
        dirs, face_idxs, normals = m._view_inputs(batch)
        target = batch["expected_rgbs"].contiguous()
        ridx = m._ray_idxs(batch)
        B = batch_rays(batch)
        scale = 1.0 / (3.0 * float(global_rays if global_rays is not None else B))
        self.step_dev.add_(1)
        pred, flat_full = torch.ops.inf_b200.field_train_step(
            table, vids, bary, self.packed, target, dirs, face_idxs, normals, m.cfg.key(), dense, scale, self.flat,
            self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev, ridx)
        flat, sums = ops.split_flat(flat_full)
        m._packed, m._packed_key = self.packed, None
        return sums, pred, flat

    def backward_only(self, batch, global_rays):
        """Multi rank, first half: gradients of this rank's rays into the flat buffer (no update)."""
        m = self.model
        table, vids, bary, dense = m._inputs(batch)
        dirs, face_idxs, normals = m._view_inputs(batch)
        target = batch["expected_rgbs"].contiguous()
        scale = 1.0 / (3.0 * float(global_rays))
        pred, flat_full = torch.ops.inf_b200.field_forward_backward(
            table, vids, bary, self.packed, target, None, dirs, face_idxs, normals, m.cfg.key(), dense, scale,
            m._ray_idxs(batch))
        flat, sums = ops.split_flat(flat_full)
        return sums, pred, flat

    def apply(self, flat_grad):
        """Multi rank, second half (after the allreduce of ``flat_grad``): Adam + operand refresh in one kernel."""
        self.step_dev.add_(1)
        torch.ops.inf_b200.field_adam_pack(flat_grad, self.packed, self.model.cfg.key(), self.flat, self.exp_avg,
                                           self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev)
        self.model._packed, self.model._packed_key = self.packed, None

    # ---- torch.optim.Adam-compatible checkpoints -------------------------------------------------------------------
    @property
    def step_count(self) -> int:
        return int(self.step_dev.item())

    def state_dict(self):
        state = {}
        for i, (off, r, c) in enumerate(self.layout):
            shape = self.model._ordered_params()[i].shape
            state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[off:off + r * c].view(shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + r * c].view(shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "params": list(range(len(self.layout)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"])
        for i, (off, r, c) in enumerate(self.layout):
            st = sd["state"].get(i)
            if st is None:
                continue
            self.exp_avg[off:off + r * c].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + r * c].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_dev.fill_(int(float(st["step"])))

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None


def fused_train_step_async(model, fopt: "FusedFieldAdam", batch, group=None):
    """One optimisation step with the fused optimiser; returns (loss device scalar, pred)."""
    world = parallel.world_size(group)
    B = batch_rays(batch)
    if world == 1:
        sums, pred, _ = fopt.fused_l1_step(batch)
    else:
        sums, pred, flat = fopt.backward_only(batch, B * world)
        parallel.allreduce_flat(flat, sums, group)
        fopt.apply(flat)
    _launched(batch)
    return sums[0] / (3.0 * B * world), pred


def fused_train_step(model, fopt: "FusedFieldAdam", batch, group=None) -> Tuple[float, torch.Tensor]:
    """trainer.py:71-84 with the fused optimiser: returns (loss.item(), pred_rgbs)."""
    loss, pred = fused_train_step_async(model, fopt, batch, group)
    return loss.item(), pred


class GraphedTrainStep:
    """``Trainer._train_step`` (trainer.py:71-84) for ``loss_type: L1`` replayed from CUDA graphs.

    One optimisation step is five kernels of ours (operand pack, forward/backward, wgrad, reduce) plus Adam; at the
    shipped ``batch_size: 4096`` the GPU work is ~0.1 ms and the host-side launch path dominates, and with a
    ``loss.item()`` after every step (trainer.py:84) the launch path is exposed at every batch size.  The whole step
    is therefore captured once per distinct set of input buffers (the staging slots of :class:`HostRayLoader`, or a
    resident batch) and replayed with a single launch.  Requirements: the optimiser is ``torch.optim.Adam(...,
    capturable=True)``; batches of one shape; single process or ``torch.distributed`` with NCCL (the gradient
    allreduce is captured too).

    ``step(batch) -> (loss.item(), pred)`` keeps the reference contract; ``step_async`` returns the device scalar.
    """

    def __init__(self, model, optim, group=None):
        self.model, self.optim, self.group = model, optim, group
        self.graphs = {}
        self._idx_bufs = {}

    @staticmethod
    def _key(batch):
        return tuple((k, v.data_ptr(), tuple(v.shape)) for k, v in sorted(batch.items()) if torch.is_tensor(v)) + \
            (id(batch.get("eigenfunction_table")),)

    # The step is split at the gradient allreduce: [pack, forward/backward, wgrad, reduce] | NCCL | [loss, Adam].
    # Single-process runs capture both halves into one graph; multi-rank runs capture them separately and issue
    # the collective eagerly between the two replays (no NCCL call inside a capture).
    def _front(self, batch):
        world = parallel.world_size(self.group)
        B = batch_rays(batch)
        if isinstance(self.optim, FusedFieldAdam):
            if world == 1:
                return self.optim.fused_l1_step(batch)                   # includes the update
            return self.optim.backward_only(batch, B * world)
        self.optim.zero_grad(set_to_none=True)
        return self.model.l1_step(batch, global_rays=B * world)          # sums, pred, flat

    def _back(self, sums, n_global, flat=None):
        loss = sums[0] / (3.0 * n_global)
        if isinstance(self.optim, FusedFieldAdam):
            if flat is not None:
                self.optim.apply(flat)
        else:
            self.optim.step()
        return loss

    def _eager(self, batch):
        world = parallel.world_size(self.group)
        sums, pred, flat = self._front(batch)
        if world > 1:
            parallel.allreduce_flat(flat, sums, self.group)
        return self._back(sums, batch_rays(batch) * world, flat if world > 1 else None), pred

    def _capture(self, batch):
        world = parallel.world_size(self.group)
        n_global = batch_rays(batch) * world
        g1 = torch.cuda.CUDAGraph()
        if world == 1:
            with torch.cuda.graph(g1):
                loss, pred = self._eager(batch)
            return (g1, None, None, None, loss, pred)
        with torch.cuda.graph(g1):
            sums, pred, flat = self._front(batch)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, pool=g1.pool()):
            loss = self._back(sums, n_global, flat)
        return (g1, g2, flat, sums, loss, pred)

    def step_async(self, batch):
        """Exactly one optimisation step per call.  The first call with a given set of input buffers runs eagerly
        (it also sizes the workspace and creates the optimiser state, which must exist before a capture: state
        tensors allocated during capture would be re-zeroed by every replay); the second call captures, and from
        then on the step is one graph launch (two around the allreduce when multi-rank)."""
        ridx = batch.get("ray_idxs")
        if ridx is not None:
            # a fused-select batch differs from the previous one only in its ray indices: stage them in a static buffer
            # so that every batch of this shape replays the same graph
            buf = self._idx_bufs.get((ridx.shape[0], ridx.dtype))
            if buf is None:
                buf = self._idx_bufs[(ridx.shape[0], ridx.dtype)] = torch.empty_like(ridx)
            buf.copy_(ridx)
            batch = dict(batch, ray_idxs=buf)
        key = self._key(batch)
        ent = self.graphs.get(key)
        if ent is None:
            self.graphs[key] = "warm"
            out = self._eager(batch)
            _launched(batch)
            return out
        if ent == "warm":
            ent = self.graphs[key] = self._capture(batch)
        g1, g2, flat, sums, loss, pred = ent
        g1.replay()
        if g2 is not None:
            parallel.allreduce_flat(flat, sums, self.group)
            g2.replay()
        _launched(batch)
        self.model.invalidate_packed()                      # the replay changed the parameters behind torch's back
        return loss, pred

    def step(self, batch) -> Tuple[float, torch.Tensor]:
        loss, pred = self.step_async(batch)
        return loss.item(), pred
