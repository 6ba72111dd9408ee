"""Two-rank run of the Trainer entry point (needs 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_trainer_check.py

Every rank trains on its half of every batch; after training the weights must be bit-identical on both ranks, the epoch
scalars (allreduced sums) equal, and only rank 0 may have written model.pt / checkpoint.pt / model_last_epoch.pt.
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402

import inf_b200  # noqa: E402
from inf_b200 import ray_store, train  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    c = load_golden("tf_k64")
    n, k = c["vids"].shape[0], c["table"].shape[1]
    tmp = tempfile.mkdtemp(prefix=f"inf_mg_r{rank}_")
    np.save(os.path.join(tmp, "efuncs.npy"), c["table"].double().numpy())
    for split, sl in (("train", slice(0, 192)), ("eval", slice(192, n))):
        with ray_store.RayCacheWriter(os.path.join(tmp, split)) as w:
            m = sl.stop - sl.start
            w.append_hits(c["vids"][sl], c["bary"][sl], c["target"][sl], torch.zeros(m, 3), torch.zeros(m, dtype=torch.int64))
    out_dir = os.path.join(tmp, "out")
    cfg = {"seed": 3, "data": {"eigenfunctions_path": os.path.join(tmp, "efuncs.npy"), "rescale_strategy": "unscaled",
                               "preproc_data_path_train": os.path.join(tmp, "train"), "preproc_data_path_eval": os.path.join(tmp, "eval"),
                               "img_height": 8, "img_width": 8},
           "model": {"k": k, "num_layers": 6, "mlp_hidden_dim": 128, "skip_layer_idx": 3},
           "training": {"out_dir": out_dir, "epochs": 3, "batch_size": 64, "lr": 1e-3, "loss_type": "L1", "print_every": 1,
                        "render_every": 1000, "checkpoint_every": 1}}
    t = train.run(cfg)
    world = dist.get_world_size()
    flat = torch.cat([p.detach().reshape(-1) for p in t.model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], g) for g in gathered), "weights diverged across ranks"
    sc = t.writer.scalars if hasattr(t.writer, "scalars") else None
    if sc is not None:
        mine = torch.tensor([v for _, v in sc["Train_Loss"]] + [v for _, v in sc["Val_Loss"]], dtype=torch.float64, device=flat.device)
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        assert all(torch.equal(allm[0], a) for a in allm), "epoch scalars differ across ranks"
        assert sc["Train_Loss"][-1][1] < sc["Train_Loss"][0][1], "it did not train"
    have = sorted(f for f in os.listdir(out_dir) if f.endswith(".pt"))
    assert have == (["checkpoint.pt", "model.pt", "model_last_epoch.pt"] if rank == 0 else []), (rank, have)
    assert inf_b200.kernel_error_word(f"cuda:{os.environ['LOCAL_RANK']}") == 0
    dist.barrier()
    if rank == 0:
        print("multi-gpu trainer check ok:", {k_: [round(v, 5) for _, v in vals] for k_, vals in (sc or {}).items()})
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
