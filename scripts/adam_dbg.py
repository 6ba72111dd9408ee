import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import inf_b200
from inf_b200 import trainer
from conftest import load_golden
import test_vertex_path as T
DEV = "cuda:0"
for path in ("vertex",):
    c = load_golden("tf_k100_l5s1")
    g = torch.Generator().manual_seed(9)
    V, B, k = 300, 512, c["k"]
    E = T._hks_like_table(V, k, g)
    params = {n: t.clone() for n, t in c["params"].items()}
    colmax = E.abs().amax(0)
    for n in ("layers.0.0.weight", f"layers.{c['skip']}.Ly.weight"):
        params[n] = params[n] / colmax[None, :] / 4
    vids = torch.randint(0, V, (B, 3), generator=g)
    u = torch.rand(B, 3, generator=g)
    bary = u / u.sum(-1, keepdim=True)
    target = torch.rand(B, 3, generator=g)
    tab = inf_b200.EigenTable.from_dense(E, device=DEV, path=path)
    b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": vids.to(DEV), "barycentric_coords": bary.to(DEV), "expected_rgbs": target.to(DEV)}
    cfg = dict(k=k, num_layers=c["num_layers"], mlp_hidden_dim=128, skip_layer_idx=c["skip"])
    ma = inf_b200.make_model(cfg); ma.load_state_dict(params); ma = ma.to(DEV)
    mb = inf_b200.make_model(cfg); mb.load_state_dict(params); mb = mb.to(DEV)
    oa = trainer.FusedFieldAdam(ma, lr=1e-5)
    ob = torch.optim.Adam(mb.parameters(), lr=1e-5)
    for i in range(2):
        la, _ = trainer.fused_train_step(ma, oa, b)
        lb, _ = trainer.train_step(mb, ob, b)
        print("step", i, la, lb)
        for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            d = (pa - pb).abs()
            j = d.argmax().item()
            ga = pa.grad.flatten()[j].item() if pa.grad is not None else None
            gb = pb.grad.flatten()[j].item() if pb.grad is not None else None
            gd = (pa.grad - pb.grad).abs().max().item() if pa.grad is not None and pb.grad is not None else None
            print(f"  {n:28s} max|dp| {d.max().item():.3e} at {j}: pa {pa.flatten()[j].item():.4e} pb {pb.flatten()[j].item():.4e} p0 {params[n].flatten()[j].item():.4e} ga {ga} gb {gb} max|dg| {gd}")
