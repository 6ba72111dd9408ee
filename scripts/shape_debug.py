"""Development aid: per-parameter gradient errors of one shape of test_shape_coverage_against_oracle on both paths."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import inf_b200
from inf_b200 import ops
from oracle import field_oracle as O
DEV = "cuda:0"
k, L, skip = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (200, 8, 6)
g = torch.Generator().manual_seed(k + L)
V, F_, B = 400, 700, 517
E = torch.randn(V, k, generator=g) / 6
normals = torch.nn.functional.normalize(torch.randn(F_, 3, generator=g), dim=-1)
fid = torch.randint(0, F_, (B,), generator=g)
dirs = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1)
vids = torch.randint(0, V, (B, 3), generator=g)
u = torch.rand(B, 3, generator=g)
bary = u / u.sum(-1, keepdim=True)
target = torch.rand(B, 3, generator=g)
params = O.make_params(k, num_layers=L, skip=skip, seed=L)
for path in ("ray", "vertex"):
    m = inf_b200.make_model(dict(k=k, num_layers=L, mlp_hidden_dim=128, skip_layer_idx=skip)); m.load_state_dict(params); m = m.to(DEV)
    tab = inf_b200.EigenTable.from_dense(E, device=DEV, path=path)
    b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": vids.to(DEV), "barycentric_coords": bary.to(DEV), "expected_rgbs": target.to(DEV)}
    sums, pred, flat = m.l1_step(b)
    torch.cuda.synchronize()
    proj = ops.project_vertices(m.cfg, tab, m._packed_weights()).float().cpu().numpy() if path == "vertex" else None
    for use_proj in ([False, True] if path == "vertex" else [False]):
        loss, y, ge = O.emulated_fp16_grads(params, E, vids, bary, target, L, skip, vertex_path=(path == "vertex"), proj=proj if use_proj else None)
        print(path, "proj" if use_proj else "", "pred err vs emu", float((pred.cpu().double() - torch.from_numpy(y)).abs().max()))
        for n, p in m.named_parameters():
            e = torch.from_numpy(ge[n]).float().reshape(p.shape)
            print("   %-28s rel %.5f" % (n, ((p.grad.cpu() - e).norm() / (e.norm() + 1e-12)).item()))
