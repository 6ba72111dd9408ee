"""Measurement helper (INF_DEBUG=8): where the MMA issuer of CTA 0 of vertex_project_kernel waits -- table panels (A), weight
chunks (B), a free accumulator (epilogue) -- against its total time.   [INF_VP_2CTA=0] python scripts/trace_project.py"""
import os, sys
os.environ["INF_DEBUG"] = os.environ.get("INF_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import inf_b200
from inf_b200 import ops
import bench
dev = torch.device("cuda")
faces = bench.torus_faces(224, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(224 * 224, 1023, dev), device=dev, path="vertex")
torch.manual_seed(0)
m = inf_b200.make_model(bench.MODEL_CFG).to(dev)
vids, bary, target = bench.make_rays(faces, 131072, dev, seed=1)
b = {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": vids, "barycentric_coords": bary, "expected_rgbs": target}
for _ in range(3):
    m.l1_step(b)
torch.cuda.synchronize()
ws = ops._workspaces[("cuda", torch.cuda.current_device())]
ws[1024:1024 + 8 * 64].zero_()
m.l1_step(b)
torch.cuda.synchronize()
tr = ws[1024 + 8 * 50:1024 + 8 * 55].view(torch.int64).cpu().tolist()
tot, ta, tb, tg, tiles = tr
if tot:
    print(f"issuer of CTA 0: {tiles} tiles, total {tot} cyc = {tot / 1.965e3:.1f} us; waits: table panels {100 * ta / tot:.0f} %, weight chunks "
          f"{100 * tb / tot:.0f} %, free accumulator {100 * tg / tot:.0f} %, issuing {100 * (tot - ta - tb - tg) / tot:.0f} %")
else:
    print("no trace (pair kernel in use? set INF_VP_2CTA=0)")
