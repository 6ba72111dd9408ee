"""Hardware multi-rank parity of the training step (needs >= 2 GPUs on one node; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/multi_gpu_step_check.py [--path ray|vertex]

SURVEY.md section 4 item 4 with the real kernels: an N-rank sharded step equals the 1-rank step on the concatenated batch.
Checked for both ways the ranks' gradients are summed -- inside the reduction kernel over NVLink peer memory
(parallel.PeerExchange / inf_train_step_comm) and with one NCCL allreduce -- and for eager and CUDA-graph steps:
  * at equal weights (first step) the summed flat gradient and both loss sums agree with the single-rank call on the whole
    batch (fp32 re-association), and the loss trajectory agrees over all steps,
  * after several Adam steps the weights are BIT-identical on all ranks, and identical between eager and CUDA-graph steps,
  * the two exchange mechanisms agree with each other (bit for bit at two ranks).
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import inf_b200  # noqa: E402
from inf_b200 import parallel, trainer  # noqa: E402
from oracle import field_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", default="vertex")
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    k, L, skip, V, Bg = 1023, 6, 3, 3000, 8192 * world
    g = torch.Generator().manual_seed(11)
    E = torch.randn(V, k, generator=g) / 6
    faces = torch.randint(0, V, (6000, 3), generator=g)
    params = O.make_params(k, num_layers=L, skip=skip, seed=2)
    table = inf_b200.EigenTable.from_dense(E, device=dev, path=args.path)
    batches = []
    for s in range(args.steps):
        vids = faces[torch.randint(0, 6000, (Bg,), generator=g)]
        u = torch.rand(Bg, 3, generator=g)
        batches.append((vids.to(dev), (u / u.sum(-1, keepdim=True)).to(dev), torch.rand(Bg, 3, generator=g).to(dev)))

    def model_and_opt(exchange):
        m = inf_b200.make_model(dict(k=k, num_layers=L, mlp_hidden_dim=128, skip_layer_idx=skip))
        m.load_state_dict(params)
        m = m.to(dev)
        o = trainer.FusedFieldAdam(m, lr=1e-3)
        if exchange:
            o.enable_peer_exchange()
        return m, o

    def batch_of(s, lo, hi):
        v, b, t = batches[s]
        return {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": v[lo:hi].contiguous(), "barycentric_coords": b[lo:hi].contiguous(),
                "expected_rgbs": t[lo:hi].contiguous()}

    lo, hi = parallel.shard_range(Bg)
    # ---- single-rank run on the whole batch (every rank computes it locally: no collective involved) ----
    m1, o1 = model_and_opt(False)
    single = []
    for s in range(args.steps):
        sums, pred, flat = o1.fused_l1_step(batch_of(s, 0, Bg))
        single.append((sums.clone(), flat.clone()))
    w1 = o1.flat.clone()
    results = {}
    for mode in ("peer", "nccl", "peer_graphed"):
        m, o = model_and_opt(mode.startswith("peer"))
        step = trainer.GraphedTrainStep(m, o) if mode == "peer_graphed" else None
        worst_g = worst_l = 0.0
        for s in range(args.steps):
            b = batch_of(s, lo, hi)
            if step is not None:
                # the graph is keyed on the input buffers: keep them fixed, refill in place
                if s == 0:
                    static = {k_: (v.clone() if torch.is_tensor(v) else v) for k_, v in b.items()}
                else:
                    for k_, v in b.items():
                        if torch.is_tensor(v):
                            static[k_].copy_(v)
                step.step_async(static)
                continue
            if o.exchange is not None:
                sums, pred, flat = o.fused_l1_step(b, global_rays=Bg)
            else:
                sums, pred, flat = o.backward_only(b, Bg)
                parallel.allreduce_flat(flat, sums)
                o.apply(flat)
            rs, rf = single[s]
            if s == 0:      # same weights on both sides: the summed gradient must agree tightly
                worst_g = ((flat - rf).norm() / rf.norm()).item()
            # later steps: Adam's first updates are +-lr whatever the gradient's size, so entries whose gradient is at the
            # noise floor take noise-determined signs and the two runs' weights drift apart in directions the loss does
            # not see -- the loss trajectory is the quantity that must agree
            worst_l = max(worst_l, ((sums - rs).abs() / rs.abs()).max().item())
        torch.cuda.synchronize()
        assert inf_b200.kernel_error_word(dev) == 0, "kernel watchdog / peer timeout fired"
        mine = o.flat.detach().clone()
        allw = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allw, mine)
        identical = all(torch.equal(allw[0], w) for w in allw)
        dw = ((mine - w1).norm() / w1.norm()).item()
        results[mode] = dict(grad_rel=worst_g, loss_rel=worst_l, ranks_bit_identical=identical, weights_vs_single_rank=dw)
        assert identical, (mode, "weights differ across ranks")
        # fp32 re-association; on the vertex path also one fp16 rounding of the per-rank vertex gradients instead of the joint one
        if mode != "peer_graphed":
            assert worst_g <= 1e-4 and worst_l <= 2e-3, (mode, worst_g, worst_l)
        results[mode]["weights"] = mine
    assert torch.equal(results["peer"]["weights"], results["peer_graphed"]["weights"]), "graphed steps differ from eager steps"
    cross = ((results["peer"]["weights"] - results["nccl"]["weights"]).norm() / w1.norm()).item()
    if world == 2:
        assert cross == 0.0, cross      # two ranks: g0 + g1 whichever mechanism adds them
    dist.barrier()
    if rank == 0:
        print(f"multi-gpu step check ok: world {world}, path {args.path}, {args.steps} steps of {Bg} rays;",
              {m_: {k_: v for k_, v in r.items() if k_ != "weights"} for m_, r in results.items()}, "peer vs nccl weights rel", cross)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
