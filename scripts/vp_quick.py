"""Quick GPU check of the vertex-projected path against the per-ray path and the oracle (development aid)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inf_b200
from inf_b200 import ops
from conftest import load_golden
from oracle import field_oracle as O

DEV = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
for name in ["tf_k64", "tf_k1023", "tf_k100_l5s1", "view_intrinsic", "view_extrinsic"]:
    c = load_golden(name)
    cfg = dict(k=c["k"], num_layers=c["num_layers"], mlp_hidden_dim=128, skip_layer_idx=c["skip"])
    mesh = None
    if c.get("view") is not None:
        import types
        cfg["view_dependence"] = c["view"]
        mesh = types.SimpleNamespace(face_normals=c["face_normals"].numpy())
    m = inf_b200.make_model(cfg, mesh=mesh); m.load_state_dict(c["params"]); m = m.to(DEV)
    for path in ("ray", "vertex"):
        tab = inf_b200.EigenTable.from_dense(c["table"], device=DEV, path=path)
        b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": c["vids"].to(DEV), "barycentric_coords": c["bary"].to(DEV),
             "expected_rgbs": c["target"].to(DEV)}
        if c.get("view") is not None:
            b["unit_ray_dirs"] = c["unit_ray_dirs"].to(DEV); b["hit_face_idxs"] = c["face_idxs"].to(DEV)
        if mode == "infer":
            with torch.no_grad():
                pred = m(b).cpu()
            torch.cuda.synchronize()
            err = (pred - c["pred"]).abs()
            print(name, path, "err word", inf_b200.kernel_error_word(DEV), "max", err.max().item(), "mean", err.mean().item(), flush=True)
        else:
            sums, pred, flat = m.l1_step(b)
            torch.cuda.synchronize()
            err = (pred.cpu() - c["pred"]).abs()
            kw = {}
            if c.get("view") is not None:
                kw = dict(view=c["view"], unit_ray_dirs=c["unit_ray_dirs"], hit_face_idxs=c["face_idxs"], face_normals=c["face_normals"])
            _, _, ge = O.emulated_fp16_grads(c["params"], c["table"], c["vids"], c["bary"], c["target"], c["num_layers"], c["skip"],
                                             vertex_path=(path == "vertex"), **kw)
            worst, worst_e = 0, 0
            for n, p in m.named_parameters():
                g = p.grad.cpu(); r = c["grads"][n]
                e = torch.from_numpy(ge[n]).float().reshape(g.shape)
                rel = ((g - r).norm() / (r.norm() + 1e-12)).item(); rele = ((g - e).norm() / (e.norm() + 1e-12)).item()
                if rele > 1e-2: print("   ", n, rel, rele)
                worst = max(worst, rel); worst_e = max(worst_e, rele)
            print(name, path, "err word", inf_b200.kernel_error_word(DEV), "rgb max", err.max().item(), "loss", sums[0].item() / (3 * pred.shape[0]),
                  c["loss"], "grad rel", worst, "emu", worst_e, flush=True)
