"""Tiny forward / training call for compute-sanitizer and debugging: python scripts/tiny_run.py [train] [view] [ray|vertex]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import inf_b200
from conftest import load_golden
train = "train" in sys.argv
name = "view_intrinsic" if "view" in sys.argv else "tf_k64"
path = "vertex" if "vertex" in sys.argv else "ray"
c = load_golden(name)
cfg = dict(k=c["k"], num_layers=c["num_layers"], mlp_hidden_dim=128, skip_layer_idx=c["skip"])
mesh = None
if c["view"] is not None:
    cfg["view_dependence"] = c["view"]
    mesh = types.SimpleNamespace(face_normals=c["face_normals"].numpy())
m = inf_b200.make_model(cfg, mesh=mesh)
m.load_state_dict(c["params"])
m = m.to("cuda")
b = {"eigenfunction_table": inf_b200.EigenTable.from_dense(c["table"], device="cuda", path=path),
     "vertex_idxs_of_hit_faces": c["vids"].cuda(), "barycentric_coords": c["bary"].cuda(), "expected_rgbs": c["target"].cuda()}
if c["view"] is not None:
    b["unit_ray_dirs"] = c["unit_ray_dirs"].cuda(); b["hit_face_idxs"] = c["face_idxs"].cuda()
if train:
    sums, pred, flat = m.l1_step(b)
else:
    with torch.no_grad():
        pred = m(b)
torch.cuda.synchronize()
print(path, "train" if train else "infer", name, "err word", hex(inf_b200.kernel_error_word("cuda")), "max err", float((pred.cpu() - c["pred"]).abs().max()))
