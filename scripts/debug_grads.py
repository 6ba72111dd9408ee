import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden
import inf_b200
DEV = "cuda:0"
for name in sys.argv[1:] or ["tf_k64", "tf_k1023"]:
    c = load_golden(name)
    m = inf_b200.make_model(dict(k=c["k"], num_layers=c["num_layers"], mlp_hidden_dim=128, skip_layer_idx=c["skip"]))
    m.load_state_dict(c["params"]); m = m.to(DEV)
    tab = inf_b200.EigenTable.from_dense(c["table"], device=DEV)
    b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": c["vids"].to(DEV), "barycentric_coords": c["bary"].to(DEV),
         "expected_rgbs": c["target"].to(DEV)}
    sums, pred, flat = m.l1_step(b)
    torch.cuda.synchronize()
    print(name, "B", c["target"].shape[0], "loss", sums[0].item() / (3 * c["target"].shape[0]), c["loss"])
    for n, p in m.named_parameters():
        g = p.grad.cpu(); r = c["grads"][n]
        rel = (g - r).norm() / r.norm()
        cos = (g * r).sum() / (g.norm() * r.norm())
        print(f"  {n:28s} rel {rel:.4f} cos {cos:.5f} norm ratio {g.norm()/r.norm():.4f}")
    g = m.layers[0][0].weight.grad.cpu(); r = c["grads"]["layers.0.0.weight"]
    # error by row (n) and by column block (k)
    er = (g - r).norm(dim=1) / (r.norm(dim=1) + 1e-12)
    print("  W0 per-row rel err: max", er.max().item(), "median", er.median().item())
    ec = (g - r).norm(dim=0) / (r.norm(dim=0) + 1e-12)
    print("  W0 per-col rel err: max", ec.max().item(), "median", ec.median().item(), "argmax", ec.argmax().item())
