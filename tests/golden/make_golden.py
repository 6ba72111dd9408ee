"""Generate golden input/output vectors by running the LIVE reference code.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the unmodified reference ``model.py`` / ``layers.py`` / ``mesh.py`` from
/root/reference (``igl``, ``trimesh``, ``torchinfo``, ``imageio`` are replaced by empty stub
modules because they are not installed and not used by the functions on the hot path),
builds the reference modules with ``make_model``, runs ``get_k_eigenfunc_vec_vals`` ->
``model(batch)`` -> ``F.l1_loss`` -> ``backward`` exactly like ``RayDataLoader.__next__``
(ray_dataloader.py:115-145) and ``Trainer._train_step`` (trainer.py:71-84), and writes every
input and output into ``tests/golden/<case>.npz``.  The tests then check the CPU oracle and
the CUDA path against these files without needing the reference.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    for name in ("igl", "trimesh", "torchinfo", "imageio", "imageio.plugins", "imageio.plugins.freeimage"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["torchinfo"].summary = lambda *a, **k: None
    sys.modules["imageio"].plugins = sys.modules["imageio.plugins"]
    sys.modules["imageio.plugins"].freeimage = sys.modules["imageio.plugins.freeimage"]
    sys.modules["imageio.plugins.freeimage"].download = lambda *a, **k: None
    sys.path.insert(0, REF)
    import model as ref_model   # noqa: E402
    import mesh as ref_mesh     # noqa: E402
    return ref_model, ref_mesh


def _rays(gen, V, F_, B):
    faces = torch.stack([torch.randperm(V, generator=gen)[:3] for _ in range(F_)])          # (F,3) distinct ids
    fid = torch.randint(0, F_, (B,), generator=gen)
    vids = faces[fid].to(torch.int64)
    u = torch.rand(B, 3, generator=gen)
    bary = (u / u.sum(-1, keepdim=True)).to(torch.float32)
    normals = F.normalize(torch.randn(F_, 3, generator=gen), dim=-1)
    dirs = F.normalize(torch.randn(B, 3, generator=gen), dim=-1)
    # flip so that -dir faces the normal (a camera ray hits the front side)
    facing = (-dirs * normals[fid]).sum(-1)
    dirs = torch.where(facing[:, None] < 0, -dirs, dirs)
    target = torch.rand(B, 3, generator=gen)
    return faces, fid.to(torch.int64), vids, bary, normals, dirs.to(torch.float32), target


CASES = {
    # name: (V, F, B, k, num_layers, skip, view_cfg)
    "tf_k64": (300, 500, 257, 64, 6, 3, None),
    "tf_k1023": (48, 80, 130, list(range(0, 256)) + list(range(1793, 2304)) + list(range(3840, 4096)), 6, 3, None),
    "tf_k100_l5s1": (90, 150, 77, 100, 5, 1, None),
    "view_intrinsic": (120, 200, 140, 100, 6, 3,
                       dict(strategy="intrinsic", in_dim_view_dir=1, include_view_dir=True, embed_size=4,
                            bottleneck_vec_dim=128, directional_hidden_dim=64)),
    "view_extrinsic": (120, 200, 96, 72, 6, 3,
                       dict(strategy="extrinsic", in_dim_view_dir=3, include_view_dir=True, embed_size=4,
                            bottleneck_vec_dim=128, directional_hidden_dim=64)),
}


def main():
    ref_model, ref_mesh = _import_reference()
    for ci, (name, (V, F_, B, k, L, skip, view)) in enumerate(CASES.items()):
        gen = torch.Generator().manual_seed(1000 + ci)
        kdim = len(k) if isinstance(k, list) else k
        table = torch.randn(V, kdim, generator=gen)
        table = (table / (table.max(0, keepdim=True).values - table.min(0, keepdim=True).values)).float()
        faces, fid, vids, bary, normals, dirs, target = _rays(gen, V, F_, B)

        cfg = dict(k=k, num_layers=L, mlp_hidden_dim=128, skip_layer_idx=skip, batchnorm=False)
        mesh = None
        if view is not None:
            cfg["view_dependence"] = view
            mesh = types.SimpleNamespace(face_normals=normals.numpy())
        torch.manual_seed(77 + ci)
        net = ref_model.make_model(cfg, mesh=mesh)

        # RayDataLoader.__next__ (ray_dataloader.py:121-143)
        x = ref_mesh.get_k_eigenfunc_vec_vals(table, vids, bary)
        batch = {"eigenfunctions": x, "expected_rgbs": target}
        if view is not None:
            batch["unit_ray_dirs"] = dirs
            batch["hit_face_idxs"] = fid
        # Trainer._train_step (trainer.py:75-81) with loss_type L1 (config.py:117-118)
        pred = net(batch)
        loss = F.l1_loss(pred, batch["expected_rgbs"])
        net.zero_grad(set_to_none=True)
        loss.backward()

        out = {
            "table": table.numpy(), "vids": vids.numpy(), "bary": bary.numpy(), "target": target.numpy(),
            "faces": faces.numpy(), "face_idxs": fid.numpy(), "face_normals": normals.numpy(),
            "unit_ray_dirs": dirs.numpy(),
            "x": x.detach().numpy(), "pred": pred.detach().numpy(), "loss": np.float32(loss.item()),
            "num_layers": np.int32(L), "skip": np.int32(skip), "k": np.int32(kdim),
        }
        names = []
        for n, prm in net.named_parameters():
            out["param/" + n] = prm.detach().numpy()
            out["grad/" + n] = prm.grad.detach().numpy()
            names.append(n)
        out["param_names"] = np.array(names)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "pred", tuple(pred.shape), "loss", float(loss), "params", len(names))

    # The reference's only self-check (ray_dataloader.py:148-205): 5 rays, bary=[1,0,0] -> x = E[vids[:,0]]
    gen = torch.Generator().manual_seed(5)
    efuncs = torch.rand(10, 5, generator=gen)
    vids = torch.tensor([[0, 1, 2], [1, 2, 3], [7, 8, 9], [5, 6, 7], [3, 4, 5]])
    bary = torch.tensor([[1, 0, 0]] * 5, dtype=torch.float32)
    x = ref_mesh.get_k_eigenfunc_vec_vals(efuncs, vids, bary)
    np.savez_compressed(os.path.join(OUT, "dataloader_selfcheck.npz"), table=efuncs.numpy(), vids=vids.numpy(),
                        bary=bary.numpy(), x=x.numpy())
    print("dataloader_selfcheck", tuple(x.shape))


if __name__ == "__main__":
    main()
