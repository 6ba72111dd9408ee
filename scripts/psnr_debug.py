"""Development aid: the fixed-step PSNR parity run of tests/test_gpu_parity.py for both paths, twice each (determinism)."""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inf_b200
from inf_b200 import ops, trainer
from conftest import VIEW_CFGS
from oracle import field_oracle as O
import test_gpu_parity as T
DEV = "cuda:0"
view_name = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] != "none" else None
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
view = VIEW_CFGS[view_name] if view_name else None
k, L, skip, B, lr = 64, 6, 3, 512, 1e-4
E, vids, bary, target, dirs, fid, normals = T._training_problem(view)
ntrain = 4096
params = O.make_params(k, num_layers=L, skip=skip, view=view, seed=5)
gi = torch.Generator().manual_seed(9)
idx_lists = [torch.randperm(ntrain, generator=gi)[:B] for _ in range(steps)]
kw = dict(num_layers=L, skip=skip)
if view is not None:
    kw.update(view=view, unit_ray_dirs=dirs, hit_face_idxs=fid, face_normals=normals)
ref_params, ref_losses = O.adam_train_steps(params, E, vids, bary, target, idx_lists, lr=lr, **kw)
hold = slice(ntrain, None)
kw_h = dict(kw)
if view is not None:
    kw_h.update(unit_ray_dirs=dirs[hold], hit_face_idxs=fid[hold])
with torch.no_grad():
    pr = O.field_forward(ref_params, E, vids[hold], bary[hold], **kw_h)
psnr_ref = O.epoch_psnr(float(((pr - target[hold]) ** 2).sum()), pr.shape[0])
print("ref psnr", psnr_ref, "ref loss", ref_losses[-1])
cfg = dict(k=k, num_layers=L, mlp_hidden_dim=128, skip_layer_idx=skip)
mesh = None
if view is not None:
    cfg["view_dependence"] = view
    mesh = types.SimpleNamespace(face_normals=normals.numpy())
dv, db, dt, dd, df = (t.to(DEV) for t in (vids, bary, target, dirs, fid))
for path in ("ray", "vertex"):
    finals = []
    for rep in range(2):
        m = inf_b200.make_model(cfg, mesh=mesh); m.load_state_dict(params); m = m.to(DEV)
        opt = torch.optim.Adam(m.parameters(), lr=lr)
        tab = inf_b200.EigenTable.from_dense(E, device=DEV, path=path)
        def batch(sel):
            b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": dv[sel].contiguous(), "barycentric_coords": db[sel].contiguous(),
                 "expected_rgbs": dt[sel].contiguous()}
            if view is not None:
                b["unit_ray_dirs"] = dd[sel].contiguous(); b["hit_face_idxs"] = df[sel].contiguous()
            return b
        losses = [trainer.train_step(m, opt, batch(idx.to(DEV)))[0] for idx in idx_lists]
        with torch.no_grad():
            pg = m(batch(hold)).cpu()
        psnr = O.epoch_psnr(float(((pg - target[hold]) ** 2).sum()), pg.shape[0])
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu()
        finals.append(flat)
        dist = max(float((p.detach().cpu() - ref_params[n]).abs().max()) for n, p in m.named_parameters())
        print(path, "rep", rep, "psnr", psnr, "dpsnr", psnr - psnr_ref, "loss", losses[-1], "max |w - w_ref|", dist, "err", inf_b200.kernel_error_word(DEV))
    print(path, "bit-identical reruns:", torch.equal(finals[0], finals[1]))
