"""Measurement helper (N ranks under torchrun, INF_DEBUG=8): globaltimer stamps of the peer-memory gradient exchange
inside the reduction kernel -- first block start to last block end -- per rank, averaged over steps."""
import os, sys
os.environ["INF_DEBUG"] = os.environ.get("INF_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import inf_b200
from inf_b200 import ops, trainer
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
faces = bench.torus_faces(224, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(224 * 224, 1023, dev), device=dev)
torch.manual_seed(0)
m = inf_b200.make_model(bench.MODEL_CFG).to(dev)
o = trainer.FusedFieldAdam(m, lr=1e-4)
o.enable_peer_exchange()
B = 131072
bs = []
for i in range(4):
    v, b, t = bench.make_rays(faces, B, dev, seed=1000 * rank + i)
    bs.append({"eigenfunction_table": table, "vertex_idxs_of_hit_faces": v, "barycentric_coords": b, "expected_rgbs": t})
for i in range(5):
    o.fused_l1_step(bs[i % 4], global_rays=B * world)
torch.cuda.synchronize()
ws = ops._workspaces[("cuda", torch.cuda.current_device())]
acc = torch.zeros(2, dtype=torch.float64)
n = 20
burst = int(os.environ.get("BURST", "1"))   # steps enqueued back to back before the stamps of the LAST one are read
for i in range(n):
    for j in range(burst):
        if j == burst - 1:
            ws[1024:1024 + 8 * 64].zero_()
        o.fused_l1_step(bs[(i * burst + j) % 4], global_rays=B * world)
    torch.cuda.synchronize()
    tr = ws[1024 + 8 * 44:1024 + 8 * 49].view(torch.int64).cpu().double()
    first = float(1 << 62) - tr[0]
    acc += torch.stack([first - tr[4], tr[2] - first])
acc /= n * 1e3
allv = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
dist.all_gather(allv, acc.to(dev))
if rank == 0:
    for q, a in enumerate(allv):
        print(f"burst {burst} rank {q}: last wgrad CTA done -> first reduce block {a[0].item():6.1f} us; first block start -> last block end (sums, exchange, Adam, operand refresh) {a[1].item():6.1f} us")
dist.destroy_process_group()
