"""Measurement helper: INF_DEBUG=8 python scripts/trace_chain.py [train] [f16] [grid=N] -- prints CTA 0's chain timeline (cycles)."""
import os, sys
os.environ["INF_DEBUG"] = os.environ.get("INF_DEBUG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import inf_b200
from inf_b200 import ops
import bench
train = "train" in sys.argv
f16 = "f16" in sys.argv
dev = torch.device("cuda")
G = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("grid=")), 100)
faces = bench.torus_faces(G, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(G * G, 1023, dev), device=dev, dtype="f16" if f16 else "f32")
torch.manual_seed(0)
m = inf_b200.make_model(dict(k=1023, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3)).to(dev)
B = 131072
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
ws[1024:1024 + 8 * 4000].zero_()
run()
torch.cuda.synchronize()
tr = ws[1024:1024 + 8 * 32].view(torch.int64).cpu().tolist()
us = lambda c: c / 1.9e3
labels = {0: "producer total", 1: "producer wait A_EMPTY", 4: "gemm1 issuer total", 5: "gemm1 wait G_EMPTY", 6: "gemm1 wait A_FULL",
          7: "gemm1 wait B1_FULL", 8: "gemm1 wait x-store read", 10: "chain issuer total", 11: "chain wait ACT_READY",
          12: "chain wait B2_FULL", 16: "epilogue total", 17: "epilogue wait G_FULL", 18: "epilogue wait MMA_DONE",
          19: "epilogue fence+barrier", 20: "epilogue wait store read",
          21: "epilogue fwd steps (incl. sync)", 22: "epilogue final step", 23: "epilogue bwd steps", 24: "epilogue view prologue"}
for k, name in labels.items():
    print(f"{name:28s} {tr[k]:10d} cyc  {us(tr[k]):8.1f} us")
