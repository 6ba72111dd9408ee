"""Measurement helper: do the gradients of ONE step of the view_2000 protocol (k = 1023, view branch, B = 4096, V = 50 176)
depend on the wgrad CTA split beyond fp32 re-association?  Run once per INF_WGRAD_BIGW value (read once per process):
    INF_WGRAD_BIGW=1.0 python scripts/grad_split_check.py out_a.pt ; INF_WGRAD_BIGW=0.85 python scripts/grad_split_check.py out_b.pt
    python scripts/grad_split_check.py out_a.pt out_b.pt        (compare)"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
if len(sys.argv) == 3:
    a, b = torch.load(sys.argv[1]), torch.load(sys.argv[2])
    for path in a:
        ga, gb = a[path], b[path]
        rel = ((ga - gb).norm() / gb.norm()).item()
        print(f"{path}: |ga - gb| / |gb| = {rel:.3e}   max abs diff {(ga - gb).abs().max().item():.3e}   (|g| max {gb.abs().max().item():.3e})")
    sys.exit(0)
import inf_b200
from conftest import VIEW_CFGS
from oracle import field_oracle as O
from psnr_protocols import PROTOCOLS, problem, index_lists
pr = PROTOCOLS["view_2000"]
view = VIEW_CFGS[pr["view"]]
E, vids, bary, target, dirs, fid, normals = problem(pr)
params = O.make_params(pr["k"], num_layers=pr["L"], skip=pr["skip"], view=view, seed=5)
cfg = dict(k=pr["k"], num_layers=pr["L"], mlp_hidden_dim=128, skip_layer_idx=pr["skip"], view_dependence=view)
out = {}
idx = index_lists(pr)[0]
for path in ("ray", "vertex"):
    m = inf_b200.make_model(cfg, mesh=types.SimpleNamespace(face_normals=normals.numpy()))
    m.load_state_dict(params)
    m = m.to("cuda:0")
    tab = inf_b200.EigenTable.from_dense(E, device="cuda:0", path=path)
    b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": vids[idx].to("cuda:0"), "barycentric_coords": bary[idx].to("cuda:0"),
         "expected_rgbs": target[idx].to("cuda:0"), "unit_ray_dirs": dirs[idx].to("cuda:0"), "hit_face_idxs": fid[idx].to("cuda:0")}
    sums, pred, flat = m.l1_step(b)
    out[path] = flat.detach().cpu().clone()
torch.save(out, sys.argv[1])
print("saved", sys.argv[1], "BIGW", os.environ.get("INF_WGRAD_BIGW"))
