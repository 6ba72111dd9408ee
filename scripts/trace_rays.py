"""Measurement helper: INF_DEBUG=8 python scripts/trace_rays.py [train] [grid=N] [B=N] -- per-role cycle accounting of CTA 0 of
field_rays_kernel (vertex path)."""
import os, sys
os.environ["INF_DEBUG"] = os.environ.get("INF_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import inf_b200
from inf_b200 import ops
import bench
train = "train" in sys.argv
dev = torch.device("cuda")
G = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("grid=")), 224)
B = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("B=")), 131072)
faces = bench.torus_faces(G, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(G * G, 1023, dev), device=dev, path="vertex")
torch.manual_seed(0)
m = inf_b200.make_model(dict(k=1023, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3)).to(dev)
vids, bary, target = bench.make_rays(faces, B, dev, seed=1)
b = {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": vids, "barycentric_coords": bary, "expected_rgbs": target}
def run():
    if train:
        m.l1_step(b)
    else:
        with torch.no_grad():
            m(b)
for _ in range(3):
    run()
torch.cuda.synchronize()
ws = ops._workspaces[("cuda", torch.cuda.current_device())]
ws[1024:1024 + 8 * 64].zero_()
run()
torch.cuda.synchronize()
tr = ws[1024:1024 + 8 * 32].view(torch.int64).cpu().tolist()
us = lambda c: c / 1.9e3
labels = {0: "chain issuer total", 1: "chain wait W_FULL", 2: "chain wait ACT_READY"}
for s in (0, 1):
    o = 8 + 12 * s
    for i, n in enumerate(["total", "wait MMA_DONE", "wait FREE_X/Y", "fence+barrier", "gather (incl. sync, publish)", "fwd steps",
                           "final step", "bwd steps"]):
        labels[o + i] = f"slot {s} {n}"
for k, name in labels.items():
    print(f"{name:40s} {tr[k]:10d} cyc  {us(tr[k]):8.1f} us")
