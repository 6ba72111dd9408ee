"""Golden files for the preprocessed-ray format, written by the LIVE reference ``MeshViewPreProcessor``.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_ray_store_golden.py

The reference class (mesh.py:431-570) is used unmodified except for two things that need packages this image
does not have: ``__init__`` (loads a mesh through trimesh) is bypassed with ``object.__new__``, and
``_ray_mesh_intersect`` (pyembree) is replaced by a deterministic fake that returns synthetic hits for the rays
the real ``create_ray_origins_and_directions`` produced.  Everything that defines the on-disk format --
``cache_single_view`` (ray generation, depth check, selection, the int32/float32 casts, mesh.py:455-523) and
``write_to_disk`` (mesh.py:525-570) -- is the reference's own code.  The five ``.npy`` files it writes are stored
byte for byte in ``ray_store.npz`` next to the per-view inputs and fake-intersector outputs, and what the
reference's ``load_preprocessed_data`` (dataset.py:12-33) reads back from them.
"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, _import_reference  # noqa: E402

FILES = ("face_idxs", "vids_of_hit_faces", "barycentric_coords", "expected_rgbs", "unit_ray_dirs")


def main():
    _, ref_mesh = _import_reference()
    import dataset as ref_dataset                                   # noqa: E402  (reference dataset.py)
    assert ref_dataset.__file__.startswith(REF)
    gen = torch.Generator().manual_seed(4242)
    H, W, V, F_ = 24, 32, 500, 900
    faces = torch.randint(0, V, (F_, 3), generator=gen)

    pre = object.__new__(ref_mesh.MeshViewPreProcessor)
    tmp = tempfile.mkdtemp()
    pre.out_dir = tmp
    for name in ("vertex_idxs_of_hit_faces", "barycentric_coords", "expected_rgbs", "unit_ray_dirs", "face_idxs"):
        setattr(pre, "cache_" + name, [])

    store = {"H": H, "W": W, "n_views": 3}
    state = {}

    def fake_intersect(ray_origins, ray_directions, return_depth=False, camCv2world=None):
        L = ray_origins.shape[0]
        hit = torch.rand(L, generator=gen) < 0.7
        hit_ray_idxs = torch.nonzero(hit).reshape(-1)
        hit_ray_idxs = hit_ray_idxs[torch.randperm(hit_ray_idxs.numel(), generator=gen)]   # intersection does not keep order
        M = hit_ray_idxs.numel()
        face_idxs = torch.randint(0, F_, (M,), generator=gen)
        vids = faces[face_idxs].to(torch.int64)
        u = torch.rand(M, 3, generator=gen, dtype=torch.float64)
        bary = (u / u.sum(-1, keepdim=True)).to(torch.float32)
        out = [vids, bary, hit_ray_idxs, face_idxs]
        if return_depth:
            depth = state["depth_masked"][hit_ray_idxs.numpy()].copy()
            bad = np.arange(M) % 11 == 3                                                  # some hits fail the depth check
            depth[bad] *= 1.5
            out.append(torch.from_numpy(depth))
        state["last"] = [t.clone() for t in out]
        return tuple(out)

    pre._ray_mesh_intersect = fake_intersect
    for v in range(3):
        ang = 0.3 * v
        R = torch.tensor([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], dtype=torch.float32)
        cam = torch.cat([R, torch.tensor([[0.1 * v], [0.2], [-3.0]])], dim=1)           # 3x4 camCv2world
        K = torch.tensor([[40.0, 0, W / 2 - 0.5], [0, 38.0, H / 2 - 0.5], [0, 0, 1]])
        mask = (torch.rand(H, W, generator=gen) < 0.6)
        img = torch.rand(H, W, 3, generator=gen)
        depth = None
        if v == 1:
            depth = (2.0 + torch.rand(H, W, generator=gen, dtype=torch.float64)).numpy()
            state["depth_masked"] = depth.reshape(-1)[mask.reshape(-1).numpy()]
        pre.cache_single_view(cam, K, mask, img, depth_check=depth)
        store[f"cam{v}"], store[f"K{v}"], store[f"mask{v}"], store[f"img{v}"] = cam.numpy(), K.numpy(), mask.numpy(), img.numpy()
        if depth is not None:
            store[f"depth{v}"] = depth
        for nm, t in zip(("vids", "bary", "hit_ray_idxs", "face_idxs", "hit_depth"), state["last"]):
            store[f"hit{v}_{nm}"] = t.numpy()
    pre.write_to_disk()

    for f in FILES:
        with open(os.path.join(tmp, f + ".npy"), "rb") as fh:
            store["file_" + f] = np.frombuffer(fh.read(), dtype=np.uint8)
    loaded = ref_dataset.load_preprocessed_data(tmp)
    for kname, t in loaded.items():
        store["loaded_" + kname] = t.numpy()
    np.savez_compressed(os.path.join(OUT, "ray_store.npz"), **store)
    print("wrote ray_store.npz:", {k: tuple(v.shape) for k, v in loaded.items()})


if __name__ == "__main__":
    main()
