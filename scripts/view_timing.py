"""Measurement helper: training-step time of the view-dependent model (BASELINE.json configs[2] shape) at 131 072 rays."""
import os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import inf_b200
from inf_b200 import trainer
import bench
dev = torch.device("cuda")
G, B = 224, 131072
faces = bench.torus_faces(G, dev)
table = inf_b200.EigenTable.from_dense(bench.make_table(G * G, 1023, dev), device=dev)
normals = torch.nn.functional.normalize(torch.randn(faces.shape[0], 3, device=dev), dim=-1)
view = dict(strategy="intrinsic", in_dim_view_dir=1, include_view_dir=True, embed_size=4, bottleneck_vec_dim=128,
            directional_hidden_dim=64)
torch.manual_seed(0)
for use_view in (False, True):
    cfg = dict(k=1023, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3)
    mesh = None
    if use_view:
        cfg["view_dependence"] = view
        mesh = types.SimpleNamespace(face_normals=normals.cpu().numpy())
    m = inf_b200.make_model(cfg, mesh=mesh).to(dev)
    fopt = trainer.FusedFieldAdam(m, lr=1e-4)
    g = torch.Generator(device=dev).manual_seed(1)
    fid = torch.randint(0, faces.shape[0], (B,), device=dev, generator=g)
    u = torch.rand(B, 3, device=dev, generator=g)
    d = torch.nn.functional.normalize(torch.randn(B, 3, device=dev, generator=g), dim=-1)
    d = torch.where(((-d * normals[fid]).sum(-1) < 0)[:, None], -d, d)
    b = {"eigenfunction_table": table, "vertex_idxs_of_hit_faces": faces[fid].contiguous(),
         "barycentric_coords": (u / u.sum(-1, keepdim=True)).contiguous(), "expected_rgbs": torch.rand(B, 3, device=dev, generator=g),
         "unit_ray_dirs": d.contiguous(), "hit_face_idxs": fid}
    for _ in range(5):
        trainer.fused_train_step_async(m, fopt, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        trainer.fused_train_step_async(m, fopt, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    with torch.no_grad():
        m.eval()
        for _ in range(3):
            m(b)
        e0.record()
        for _ in range(20):
            m(b)
        e1.record()
        torch.cuda.synchronize()
    print(f"view={use_view}: train {ms:.3f} ms/step = {B / ms / 1e3:.1f} M rays/s; infer {e0.elapsed_time(e1) / 20:.3f} ms = "
          f"{B / (e0.elapsed_time(e1) / 20) / 1e3:.1f} M rays/s; err word {inf_b200.kernel_error_word(dev)}")
