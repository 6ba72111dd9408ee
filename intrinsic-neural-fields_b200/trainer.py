"""Training-step surface of the reference's ``Trainer`` (trainer.py:71-84, 158-187) on the fused path.

``train_step`` keeps ``_train_step``'s contract -- ``(loss.item(), pred_rgbs)`` after forward, loss,
zero_grad, backward and ``optim.step()`` -- and uses the fused forward+backward kernel when the loss is
the config's L1 (config.py:117-118); any other loss callable goes through autograd into the same
backward kernels.  ``Trainer`` is the reference's class around it (epoch loop, best model, scheduler, checkpoints).
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
        loss = loss_fn(pred, _targets(batch))      # the batch's own rows (the whole split's array when ray_idxs selects)
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
        self.exchange = None      # parallel.PeerExchange: the multi-GPU gradient sum inside the reduction kernel
        self._col_exp = None      # column exponents (EigenTable.col_exp) the operand images were packed for
        # the operand images are refreshed by the Adam kernel only: weights written from outside (a checkpoint restored
        # after the optimiser was built, trainer.py:214-231) must be re-packed before the next forward
        model.register_load_state_dict_post_hook(lambda module, incompatible_keys: self.repack())

    def enable_peer_exchange(self, group=None):
        """Multi-GPU (one node): sum the ranks' gradients inside the reduction kernel over NVLink peer memory instead of an
        NCCL allreduce followed by a separate optimiser launch.  Call once after ``init_process_group`` on every rank."""
        import os
        if parallel.world_size(group) > 1 and self.exchange is None and os.environ.get("INF_PEER_EXCHANGE", "1") != "0":
            self.exchange = parallel.PeerExchange(self.model.cfg, self.flat.device, group)
        return self.exchange

    def repack(self, col_exp="same"):
        """Rebuild the fp16 operand images from the flat fp32 parameters (in place: captured graphs keep their pointer)."""
        if col_exp != "same":
            self._col_exp = col_exp
        ops.pack_weights(self.model.cfg, [p.detach() for p in self.model._ordered_params()], self.packed, self._col_exp)
        self.model._packed, self.model._packed_key = self.packed, None

    def _for_table(self, batch):
        """The operand images are specific to the table's column exponents: re-pack when a batch brings another table."""
        ce = self.model._table_extras(batch)[2]
        if (None if ce is None else ce.data_ptr()) != (None if self._col_exp is None else self._col_exp.data_ptr()):
            self.repack(ce)
        return ce

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
        self._for_table(batch)
        # step_dev holds the number of completed updates; the reduction kernel uses t = step_dev + 1 and advances it
        pred, flat_full = torch.ops.inf_b200.field_train_step(
            table, vids, bary, self.packed, target, dirs, face_idxs, normals, m.cfg.key(), dense, scale, self.flat,
            self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev, ridx,
            *m._table_extras(batch), self.exchange.id if self.exchange is not None else 0)
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
        self._for_table(batch)
        pred, flat_full = torch.ops.inf_b200.field_forward_backward(
            table, vids, bary, self.packed, target, None, dirs, face_idxs, normals, m.cfg.key(), dense, scale,
            m._ray_idxs(batch), *m._table_extras(batch))
        flat, sums = ops.split_flat(flat_full)
        return sums, pred, flat

    def apply(self, flat_grad):
        """Multi rank, second half (after the allreduce of ``flat_grad``): Adam + operand refresh in one kernel."""
        self.step_dev.add_(1)
        torch.ops.inf_b200.field_adam_pack(flat_grad, self.packed, self.model.cfg.key(), self.flat, self.exp_avg,
                                           self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev,
                                           self._col_exp)
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
    elif fopt.exchange is not None:
        sums, pred, _ = fopt.fused_l1_step(batch, global_rays=B * world)      # the sum over ranks runs inside the step
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
            if self.optim.exchange is not None:
                return self.optim.fused_l1_step(batch, global_rays=B * world)   # includes the sum over ranks and the update
            return self.optim.backward_only(batch, B * world)
        self.optim.zero_grad(set_to_none=True)
        return self.model.l1_step(batch, global_rays=B * world)          # sums, pred, flat

    def _back(self, sums, n_global, flat=None):
        loss = sums[2]          # mean L1 loss = sum |d| / (3 n_global), written by the reduction kernel (no extra launch)
        if isinstance(self.optim, FusedFieldAdam):
            if flat is not None:
                self.optim.apply(flat)
        else:
            self.optim.step()
        return loss

    def _peer(self):
        return isinstance(self.optim, FusedFieldAdam) and self.optim.exchange is not None

    def _eager(self, batch):
        world = parallel.world_size(self.group)
        sums, pred, flat = self._front(batch)
        nccl = world > 1 and not self._peer()
        if nccl:
            parallel.allreduce_flat(flat, sums, self.group)
        return self._back(sums, batch_rays(batch) * world, flat if nccl else None), pred

    def _capture(self, batch):
        world = parallel.world_size(self.group)
        n_global = batch_rays(batch) * world
        g1 = torch.cuda.CUDAGraph()
        if world == 1 or self._peer():     # no NCCL call in the step: one graph at any number of ranks
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

    def step_logged(self, batch):
        """One optimisation step whose loss is read back WITHOUT draining the GPU: the scalar of this step is copied to a
        pinned host slot behind the step's kernels (4 bytes on a copy stream) and the call returns the previous step's loss
        -- ``loss.item()`` after every launch (trainer.py:84) leaves the device idle for a host round trip per step, a
        training log does not need the value before the next step is enqueued.  Returns ``(loss of the previous call or
        None, pred)``; :meth:`flush_logged` returns the last one."""
        if not hasattr(self, "_log"):
            dev = next(self.model.parameters()).device
            self._log = {"host": torch.zeros(2, dtype=torch.float32).pin_memory(), "dev": torch.zeros(2, dtype=torch.float32, device=dev),
                         "ev": [torch.cuda.Event(), torch.cuda.Event()], "stream": torch.cuda.Stream(dev), "n": 0,
                         "done": [torch.cuda.Event(), torch.cuda.Event()]}
        lg = self._log
        loss, pred = self.step_async(batch)
        cur = torch.cuda.current_stream(loss.device)
        slot = lg["n"] & 1
        # The graph's loss tensor is overwritten by the next replay, so the scalar is first copied (4 bytes, device to device,
        # in stream order) into a two-slot ring; the host copy then runs on its own stream and the next step never waits for it
        # (a slot is reused two steps later: its previous host copy has long finished).
        if lg["n"] >= 2:
            cur.wait_event(lg["ev"][slot])
        lg["dev"][slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        lg["done"][slot].record(cur)
        with torch.cuda.stream(lg["stream"]):
            lg["stream"].wait_event(lg["done"][slot])
            lg["host"][slot:slot + 1].copy_(lg["dev"][slot:slot + 1], non_blocking=True)
            lg["ev"][slot].record(lg["stream"])
        prev = None
        if lg["n"] > 0:
            lg["ev"][slot ^ 1].synchronize()        # the PREVIOUS step's 4 bytes: long done, the current step keeps running
            prev = float(lg["host"][slot ^ 1])
        lg["n"] += 1
        return prev, pred

    def flush_logged(self):
        lg = getattr(self, "_log", None)
        if lg is None or lg["n"] == 0:
            return None
        slot = (lg["n"] - 1) & 1
        lg["ev"][slot].synchronize()
        return float(lg["host"][slot])


class _ScalarLog:
    """Stand-in for ``tensorboardX.SummaryWriter`` (not installed in this image): keeps the scalars in memory."""

    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, global_step=None):
        self.scalars.setdefault(tag, []).append((global_step, float(value)))

    def add_image(self, *a, **k):
        pass


class Trainer:
    """The reference's ``Trainer`` (trainer.py:18-340) over the fused path: same constructor arguments, ``_train_step``,
    ``_eval_step``, ``evaluate``, ``test``, ``train`` (epoch loop, best-model tracking into ``model.pt``, ReduceLROnPlateau,
    ``checkpoint.pt`` save / restore with the RNG states, ``model_last_epoch.pt``).

    What differs on purpose: the epoch loop does not read ``loss.item()`` after every step (trainer.py:84,240-242) -- the
    per-batch loss and squared-error sums stay on the device and are read once per epoch; with the config's L1 loss the
    step is the fused kernel sequence, replayed from a CUDA graph when the optimiser allows it (``FusedFieldAdam`` or a
    capturable ``torch.optim`` optimiser).  Rendering validation views into TensorBoard (trainer.py:107-156) needs the
    reference's file readers and ray caster and is left to an optional ``vis_fn(trainer, epoch)`` callback; scalars go to
    ``tensorboardX`` when it is importable, else to an in-memory log (``trainer.writer.scalars``).
    """

    def __init__(self, model, optim, loss_fn, renderer, data, mesh, config, device, *, writer=None, vis_fn=None, group=None):
        import copy
        import os
        self.model, self.optim, self.loss_fn, self.renderer, self.mesh, self.config = model, optim, loss_fn, renderer, mesh, config
        tr = config["training"]
        self.use_lr_scheduler = tr.get("use_lr_scheduler", False)
        self.lr_scheduler = None
        if self.use_lr_scheduler:
            if isinstance(optim, FusedFieldAdam):
                raise NotImplementedError("use_lr_scheduler needs a torch.optim optimiser (FusedFieldAdam has a fixed lr)")
            self.lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optim, mode="min", factor=0.2)   # trainer.py:36
        self.train_data_loader, self.val_data_loader = data["train"], data["val"]
        self.test_data_loader = data.get("test", None)
        # the view branch reads face_normals[hit_face_idxs] raw: range-check every split once, where it is bound
        normals = getattr(model, "face_normals", None)
        if normals is not None:
            for dl in (self.train_data_loader, self.val_data_loader, self.test_data_loader):
                arrays = getattr(dl, "arrays", None) or getattr(dl, "host", None) or {}
                if dl is not None and arrays.get("hit_face_idxs") is not None:
                    ops.validate_indices(dl.table, None, arrays["hit_face_idxs"], int(normals.shape[0]))
        self.out_dir = tr["out_dir"]
        os.makedirs(os.path.join(self.out_dir, "logs"), exist_ok=True)
        if writer is None:
            try:
                from tensorboardX import SummaryWriter
                writer = SummaryWriter(os.path.join(self.out_dir, "logs"))
            except ImportError:
                writer = _ScalarLog()
        self.writer, self.vis_fn, self.group = writer, vis_fn, group
        self.render_every, self.print_every, self.epochs = tr.get("render_every", 1 << 30), tr.get("print_every", 1), tr["epochs"]
        self.checkpoint_every = tr.get("checkpoint_every")
        self.checkpoint_path = os.path.join(self.out_dir, "checkpoint.pt")
        self.device = device
        self.best_model_weights_path = os.path.join(self.out_dir, "model.pt")
        self.best_model = copy.deepcopy(model)                                                           # trainer.py:67
        self.model_last_epoch_path = os.path.join(self.out_dir, "model_last_epoch.pt")
        self._writes_files = parallel.rank(group) == 0       # weights are identical on every rank: rank 0 keeps the files
        self._graphed = None
        if loss_fn is F.l1_loss and (isinstance(optim, FusedFieldAdam) or
                                     any(g.get("capturable", False) for g in getattr(optim, "param_groups", []))):
            self._graphed = GraphedTrainStep(model, optim, group)

    # ---- steps ------------------------------------------------------------------------------------------------------
    def _train_step_async(self, batch):
        if self._graphed is not None:
            return self._graphed.step_async(batch)
        if isinstance(self.optim, FusedFieldAdam):
            return fused_train_step_async(self.model, self.optim, batch, self.group)
        return train_step_async(self.model, self.optim, batch, self.loss_fn, self.group)

    def _train_step(self, batch):
        """trainer.py:71-84: ``(loss.item(), pred_rgbs)``."""
        loss, pred = self._train_step_async(batch)
        return loss.item(), pred

    @torch.no_grad()
    def _eval_step(self, model, batch):
        pred = model(batch)
        return self.loss_fn(pred, _targets(batch)), pred

    def _epoch_metrics(self, model, loader, step):
        """Rays-weighted mean loss and sum of squared errors over a loader, accumulated on the device."""
        acc = torch.zeros(2, dtype=torch.float64, device=self.device)
        total = 0
        for batch in loader:
            loss, pred = step(model, batch)
            gt = _targets(batch)
            n = gt.size()[0]
            acc[0] += loss.detach().double() * n
            acc[1] += F.mse_loss(pred.detach(), gt, reduction="sum").double()
            total += n
        world = parallel.world_size(self.group)
        if world > 1:                                        # every rank saw its 1/W of the rays: one collective per epoch
            import torch.distributed as dist
            cnt = torch.tensor([float(total)], dtype=torch.float64, device=self.device)
            both = torch.cat([acc, cnt])
            dist.all_reduce(both, op=dist.ReduceOp.SUM, group=self.group)
            acc, total = both[:2], int(both[2].item())
        a = acc.tolist()                                     # the epoch's only host read
        return a[0] / total, epoch_psnr(a[1] / total)

    def evaluate(self, epoch=None):
        """trainer.py:164-187."""
        self.model.eval()
        val_loss, val_psnr = self._epoch_metrics(self.model, self.val_data_loader, self._eval_step)
        self.writer.add_scalar("Val_Loss", val_loss, epoch)
        self.writer.add_scalar("Val Epoch-PSNR", val_psnr, epoch)
        return val_loss, val_psnr

    def test(self):
        """trainer.py:189-212."""
        if self.test_data_loader is None:
            return None
        self.best_model.eval()
        test_loss, _ = self._epoch_metrics(self.best_model, self.test_data_loader, self._eval_step)
        self.writer.add_scalar("Test Loss", test_loss)
        print(f"Test Loss: {test_loss}")
        return test_loss

    # ---- checkpoints ------------------------------------------------------------------------------------------------
    def _checkpoint(self, epoch):
        import random
        import numpy as np
        return {"epoch": epoch, "model_state_dict": self.model.state_dict(), "optimizer_state_dict": self.optim.state_dict(),
                "pytorch_random_state": torch.random.get_rng_state(), "python_random_state": random.getstate(),
                "numpy_random_state": np.random.get_state()}

    def _init_or_load_checkpoint(self):
        """trainer.py:214-231."""
        import os
        import random
        import numpy as np
        if self.checkpoint_every is None or not os.path.exists(self.checkpoint_path):
            return 0
        print("Restoring from checkpoint...")
        ck = torch.load(self.checkpoint_path, map_location="cpu", weights_only=False)
        self.model.load_state_dict(ck["model_state_dict"])
        self.optim.load_state_dict(ck["optimizer_state_dict"])
        if hasattr(self.model, "invalidate_packed"):
            self.model.invalidate_packed()
        torch.random.set_rng_state(ck["pytorch_random_state"])
        random.setstate(ck["python_random_state"])
        np.random.set_state(ck["numpy_random_state"])
        print("Done.")
        return ck["epoch"] + 1

    # ---- epoch loop -------------------------------------------------------------------------------------------------
    def train(self):
        """trainer.py:233-340."""
        import copy
        import os
        import time
        print("Starting training...")
        first = self._init_or_load_checkpoint()
        min_val_loss = 1.0
        for epoch in range(first, self.epochs):
            self.model.train()
            t0 = time.time()
            train_loss, train_psnr = self._epoch_metrics(self.model, self.train_data_loader,
                                                         lambda m, b: self._train_step_async(b))
            t1 = time.time()
            self.writer.add_scalar("Train_Loss", train_loss, epoch)
            self.writer.add_scalar("Train Epoch-PSNR", train_psnr, epoch)
            val_loss, val_psnr = self.evaluate(epoch)
            if val_loss < min_val_loss:
                min_val_loss = val_loss
                if self._writes_files:
                    torch.save(self.model.state_dict(), self.best_model_weights_path)
                self.best_model = copy.deepcopy(self.model)
            if self.lr_scheduler is not None:
                self.lr_scheduler.step(val_loss)
            if epoch == 0 or (epoch + 1) % self.print_every == 0:
                print(f"Epoch: {epoch + 1} / {self.epochs}, Train Loss: {train_loss}, Train PSNR: {train_psnr}, "
                      f"Val Loss: {val_loss}, Val PSNR: {val_psnr} Epoch Time: {t1 - t0}s")
            if self.vis_fn is not None and (epoch == 0 or (epoch + 1) % self.render_every == 0):
                self.model.eval()
                self.vis_fn(self, epoch)
            if self._writes_files and self.checkpoint_every is not None and epoch % self.checkpoint_every == 0:
                torch.save(self._checkpoint(epoch), self.checkpoint_path)
            if self._writes_files and epoch > 0 and (epoch + 1) == 200:                      # trainer.py:315-334
                torch.save(self._checkpoint(epoch), os.path.join(self.out_dir, f"checkpoint_{epoch}.pt"))
                torch.save(self.best_model.state_dict(), os.path.join(self.out_dir, f"best_model_checkpoint_{epoch}.pt"))
        self.test()
        print("Done.")
        if self._writes_files:
            torch.save(self.model.state_dict(), self.model_last_epoch_path)


def _targets(batch):
    """The batch's target rows (the whole split's array when the batch selection is folded into the kernels)."""
    r = batch.get("ray_idxs")
    return batch["expected_rgbs"] if r is None else batch["expected_rgbs"][r.long()]
