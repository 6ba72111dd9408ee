"""``train.py`` of the reference (train.py:22-64) on the fused path.

    python -m inf_b200.train CONFIG.yaml [--allow_checkpoint_loading]            # one GPU
    torchrun --nproc-per-node N -m inf_b200.train CONFIG.yaml                     # replaces --data_parallel (train.py:46-48)

Seeding (train.py:27-30,55-57), data (config.py:58-103), model + Adam(lr) (config.py:102-110), loss (config.py:113-122) and the
``Trainer`` epoch loop are the reference's; mesh loading (trimesh) is needed only for the view-direction branch -- pass an
object with ``face_normals`` to :func:`run`, or install trimesh for the command line.
"""
from __future__ import annotations

import argparse
import os
import random

import numpy as np
import torch


def _seed(seed: int):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def run(config: dict, mesh=None, allow_checkpoint_loading: bool = False, fused_adam: bool = True):
    """Train ``config`` on this process's GPU (rank-sharded when launched under torchrun); returns the ``Trainer``."""
    import torch.distributed as dist
    from . import trainer
    from .model import make_model
    from .ray_dataloader import get_data
    out_dir = config["training"]["out_dir"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("the B200 path has no CPU fallback: a CUDA device is required")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(device))
    # config.py:31-33, decided by rank 0 and shared: every rank leaves together instead of waiting in a collective
    taken = torch.tensor([int(int(os.environ.get("RANK", "0")) == 0 and os.path.exists(out_dir) and not allow_checkpoint_loading
                              and bool(os.listdir(out_dir)))], device=device)
    if world > 1:
        dist.broadcast(taken, src=0)
    if int(taken.item()):
        raise RuntimeError(f"out_dir '{out_dir}' exists. Exit to not overwrite old results.")
    seed = config.get("seed", 0)                                                                             # config.py:48-49
    _seed(seed)
    data = get_data(config, device)
    model = make_model(config["model"], mesh=mesh).to(device)
    lr = config["training"]["lr"]
    loss_fn = trainer.get_loss_fn(config["training"]["loss_type"])
    use_fused = fused_adam and loss_fn is torch.nn.functional.l1_loss and not config["training"].get("use_lr_scheduler", False)
    optim = trainer.FusedFieldAdam(model, lr=lr) if use_fused else torch.optim.Adam(model.parameters(), lr=lr)
    if use_fused and world > 1:
        optim.enable_peer_exchange()      # gradient sum over the node's GPUs inside the reduction kernel (no NCCL per step)
    _seed(seed)                                                    # again: the model's init consumed random numbers (train.py:55-57)
    t = trainer.Trainer(model, optim, loss_fn, None, data, mesh, config, device)
    t.train()
    return t


def main(argv=None):
    import yaml
    ap = argparse.ArgumentParser()
    ap.add_argument("config_path", type=str)
    ap.add_argument("--allow_checkpoint_loading", default=False, action="store_true")
    args = ap.parse_args(argv)
    with open(args.config_path, "r") as f:
        config = yaml.safe_load(f)
    mesh = None
    if config["model"].get("view_dependence") is not None:
        import trimesh                                             # only the view branch needs the mesh (its face normals)
        mesh = trimesh.load(config["data"]["mesh_path"], process=False)
    run(config, mesh=mesh, allow_checkpoint_loading=args.allow_checkpoint_loading)


if __name__ == "__main__":
    main()
