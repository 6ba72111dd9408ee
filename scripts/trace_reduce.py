"""Measurement helper (INF_DEBUG=8): globaltimer stamps around field_reduce_kernel inside a fused training step -- last wgrad
CTA done, first reduce block started, last block started, last block finished."""
import os, sys
os.environ["INF_DEBUG"] = os.environ.get("INF_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import inf_b200
from inf_b200 import ops, trainer
import bench
dev = torch.device("cuda")
faces = bench.torus_faces(224, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(224 * 224, 1023, dev), device=dev)
torch.manual_seed(0)
m = inf_b200.make_model(bench.MODEL_CFG).to(dev)
o = trainer.FusedFieldAdam(m, lr=1e-4)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
v, b, t = bench.make_rays(faces, B, dev, seed=7)
sb = {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": v, "barycentric_coords": b, "expected_rgbs": t}
for _ in range(5): trainer.fused_train_step_async(m, o, sb)
torch.cuda.synchronize()
ws = ops._workspaces[("cuda", torch.cuda.current_device())]
acc = torch.zeros(4, dtype=torch.float64)
n = 20
for _ in range(n):
    ws[1024:1024 + 8 * 64].zero_()
    trainer.fused_train_step_async(m, o, sb)
    torch.cuda.synchronize()
    tr = ws[1024 + 8 * 44:1024 + 8 * 49].view(torch.int64).cpu().double()
    first = float(1 << 62) - tr[0]
    acc += torch.stack([first - tr[4], tr[1] - first, tr[2] - first, tr[2] - first])
acc /= n * 1e3
print(f"B={B}: wgrad end -> first reduce block {acc[0]:.1f} us; last block start +{acc[1]:.1f}; last block end +{acc[3]:.1f} us")
