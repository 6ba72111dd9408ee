"""Measurement helper (N ranks under torchrun, INF_DEBUG=8): globaltimer stamps of the peer-memory gradient exchange
inside field_reduce_kernel -- kernel start, flag published, all peers seen, end -- per rank, averaged over steps."""
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
acc = torch.zeros(3, dtype=torch.float64)
n = 20
for i in range(n):
    o.fused_l1_step(bs[i % 4], global_rays=B * world)
    torch.cuda.synchronize()
    tr = ws[1024 + 8 * 40:1024 + 8 * 44].view(torch.int64).cpu().double()
    acc += torch.stack([tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2]])
acc /= n * 1e3
allv = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
dist.all_gather(allv, acc.to(dev))
if rank == 0:
    for q, a in enumerate(allv):
        print(f"rank {q}: partial sums + publish {a[0].item():7.1f} us   wait for peers {a[1].item():7.1f} us   sum over ranks + Adam + pack {a[2].item():7.1f} us")
dist.destroy_process_group()
