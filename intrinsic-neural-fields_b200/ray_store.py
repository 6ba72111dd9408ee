"""The preprocessed-ray format on either side of the hot path.

The reference caches every hit of every training view as five ``.npy`` files (``vids_of_hit_faces`` (N,3) i32,
``barycentric_coords`` (N,3) f32, ``expected_rgbs`` (N,3) f32, ``unit_ray_dirs`` (N,3) f32, ``face_idxs`` (N,) i32;
mesh.py:505-570) and reads them back widening the indices to int64 (dataset.py:12-33).  Its writer keeps one
Python tensor object per hit until ``write_to_disk`` (mesh.py:518-523): O(hits) objects and twice the data in RAM.

* ``RayCacheWriter`` is the streaming replacement of ``MeshViewPreProcessor``'s cache (mesh.py:431-570): each
  view's hits are appended to the five files as they arrive; ``write_to_disk`` only patches the headers.  The
  files are byte-identical to what the reference writes (tests/test_ray_store.py checks this against files
  produced by the live reference).
* ``load_preprocessed_rays`` is ``load_preprocessed_data`` without the int64 widening: the kernels read the int32
  ids as stored, so the arrays can go straight into pinned memory / onto the device (``RayDataLoader`` /
  ``HostRayLoader`` take either width).

Ray casting itself (trimesh + pyembree, mesh.py:208-245) stays the reference's CPU preprocessing: the writer takes
the intersector as a callable with ``ray_mesh_intersect``'s return convention.
"""
from __future__ import annotations

import io
import os
from typing import Callable, Dict, Optional

import numpy as np
import torch

# file stem -> (dtype, trailing shape); names and dtypes of mesh.py:505-570
FILES = {
    "face_idxs": (np.int32, ()),
    "vids_of_hit_faces": (np.int32, (3,)),
    "barycentric_coords": (np.float32, (3,)),
    "expected_rgbs": (np.float32, (3,)),
    "unit_ray_dirs": (np.float32, (3,)),
}
_HEADER_RESERVE = 128            # what np.save emits for every (N,) / (N,3) shape with fewer than ~50 digits


def _npy_header(dtype, shape) -> bytes:
    buf = io.BytesIO()
    np.lib.format.write_array_header_1_0(buf, {"descr": np.lib.format.dtype_to_descr(np.dtype(dtype)),
                                               "fortran_order": False, "shape": tuple(shape)})
    return buf.getvalue()


class NpyAppender:
    """A ``.npy`` file that grows along axis 0.  The header is reserved up front and patched on ``close``."""

    def __init__(self, path: str, dtype, tail_shape=()):
        self.path, self.dtype, self.tail = path, np.dtype(dtype), tuple(tail_shape)
        self.rows = 0
        self._fh = open(path, "wb")
        self._fh.write(b"\0" * _HEADER_RESERVE)

    def append(self, arr) -> int:
        a = np.ascontiguousarray(arr.detach().cpu().numpy() if isinstance(arr, torch.Tensor) else arr)
        if a.dtype != self.dtype:
            raise TypeError(f"{os.path.basename(self.path)}: expected {self.dtype}, got {a.dtype}")
        if a.shape[1:] != self.tail:
            raise ValueError(f"{os.path.basename(self.path)}: expected (*,{self.tail}), got {a.shape}")
        if a.size:
            self._fh.write(a.data)
        self.rows += a.shape[0]
        return a.shape[0]

    def close(self):
        if self._fh is None:
            return
        header = _npy_header(self.dtype, (self.rows,) + self.tail)
        if len(header) == _HEADER_RESERVE:
            self._fh.seek(0)
            self._fh.write(header)
            self._fh.close()
        else:                                        # astronomically long shape string: rewrite behind the real header
            self._fh.close()
            tmp = self.path + ".tmp"
            with open(self.path, "rb") as src, open(tmp, "wb") as dst:
                src.seek(_HEADER_RESERVE)
                dst.write(header)
                while True:
                    chunk = src.read(1 << 24)
                    if not chunk:
                        break
                    dst.write(chunk)
            os.replace(tmp, self.path)
        self._fh = None


def pixel_rays(camCv2world: torch.Tensor, K: torch.Tensor, mask_1d: torch.Tensor, *, H: int, W: int,
               pixel_coords: Optional[torch.Tensor] = None):
    """Origins and unit directions of the rays through the masked pixels (create_ray_origins_and_directions,
    mesh.py:171-205): ``dir = R (K^-1 [u v 1]^T)`` normalised, origin = camera centre.  ``pixel_coords`` (L,2) f32
    replaces the integer pixel grid when the caller has undistorted the pixels already (mesh.py:187-193)."""
    mask_1d = torch.as_tensor(mask_1d).reshape(-1)
    if pixel_coords is None:
        pix = torch.nonzero(mask_1d).reshape(-1)
        pixel_coords = torch.stack([pix % W, pix // W], dim=-1)                      # (u, v) = (column, row)
    L = pixel_coords.shape[0]
    homog = torch.cat([pixel_coords, torch.ones((L, 1))], dim=-1)                    # promotes to f32 like the reference
    dirs = camCv2world[:3, :3].matmul(K[:3, :3].inverse().matmul(homog.T)).T
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    return camCv2world[:, 3].unsqueeze(0).expand(L, -1), dirs


class RayCacheWriter:
    """Streaming writer of the preprocessed-ray directory (MeshViewPreProcessor, mesh.py:431-570).

    ``intersect(ray_origins, unit_ray_dirs, return_depth=False, camCv2world=None)`` must return what
    ``ray_mesh_intersect`` (mesh.py:208-245) returns: ``(vids (M,3) i64, bary (M,3) f32, hit_ray_idxs (M,) i64,
    face_idxs (M,) i64[, hit_depth (M,)])``.
    """

    def __init__(self, out_directory: str, intersect: Optional[Callable] = None):
        self.out_dir = out_directory
        self.intersect = intersect
        os.makedirs(self.out_dir, exist_ok=True)
        self._files = {stem: NpyAppender(os.path.join(self.out_dir, stem + ".npy"), dt, tail)
                       for stem, (dt, tail) in FILES.items()}
        self.num_hits = 0

    def append_hits(self, vertex_idxs_of_hit_faces, barycentric_coords, expected_rgbs, unit_ray_dirs, face_idxs) -> int:
        """Append one view's hits; applies the narrowing casts of mesh.py:505-516 (and their range checks)."""
        vids, fids = torch.as_tensor(vertex_idxs_of_hit_faces), torch.as_tensor(face_idxs)
        n = vids.shape[0]
        for name, t in (("barycentric_coords", barycentric_coords), ("expected_rgbs", expected_rgbs),
                        ("unit_ray_dirs", unit_ray_dirs), ("face_idxs", fids)):
            if t.shape[0] != n:
                raise ValueError(f"{name} has {t.shape[0]} rows, vertex ids have {n}")
        if n and (int(vids.max()) > 2**31 - 1 or int(fids.max()) > 2**31 - 1 or int(vids.min()) < 0 or int(fids.min()) < 0):
            raise OverflowError("vertex / face ids do not fit the int32 on-disk format")
        self._files["face_idxs"].append(fids.to(torch.int32))
        self._files["vids_of_hit_faces"].append(vids.to(torch.int32))
        self._files["barycentric_coords"].append(torch.as_tensor(barycentric_coords).to(torch.float32))
        self._files["expected_rgbs"].append(torch.as_tensor(expected_rgbs).to(torch.float32))
        self._files["unit_ray_dirs"].append(torch.as_tensor(unit_ray_dirs).to(torch.float32))
        self.num_hits += n
        return n

    def cache_single_view(self, camCv2world, K, mask, img, depth_check=None, pixel_coords=None) -> int:
        """One training view -> its hits appended (mesh.py:455-523).  Returns the number of hits kept."""
        if self.intersect is None:
            raise RuntimeError("RayCacheWriter needs an intersector callable to cast a view")
        H, W = mask.shape
        mask_1d = torch.as_tensor(mask).reshape(-1)
        rgbs = torch.as_tensor(img).reshape(H * W, -1)[mask_1d]
        origins, dirs = pixel_rays(camCv2world, K, mask_1d, H=H, W=W, pixel_coords=pixel_coords)
        if depth_check is None:
            vids, bary, hit_ray_idxs, face_idxs = self.intersect(origins, dirs)
        else:
            vids, bary, hit_ray_idxs, face_idxs, hit_depth = self.intersect(origins, dirs, return_depth=True,
                                                                            camCv2world=camCv2world)
            # keep hits whose depth agrees with the supplied depth map to 1 % of its mean (mesh.py:483-499)
            assert depth_check.shape == (H, W)
            want = np.asarray(depth_check).reshape(-1)[mask_1d.numpy()][hit_ray_idxs.numpy()]
            err = np.abs(np.asarray(hit_depth) - want)
            inlier = err < np.mean(want) * 1e-2
            vids, bary, hit_ray_idxs, face_idxs = vids[inlier], bary[inlier], hit_ray_idxs[inlier], face_idxs[inlier]
        return self.append_hits(vids, bary, rgbs[hit_ray_idxs], dirs[hit_ray_idxs], face_idxs)

    def write_to_disk(self):
        """Finish the five files (the data is on disk already; this patches the headers)."""
        for f in self._files.values():
            f.close()

    close = write_to_disk

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.write_to_disk()
        return False


def load_preprocessed_rays(preproc_data_path: str, *, index_dtype: torch.dtype = torch.int32, pin: bool = False,
                           device=None, mmap: bool = False) -> Dict[str, torch.Tensor]:
    """``load_preprocessed_data`` (dataset.py:12-33) with the reference's keys.  Indices stay int32 as stored
    unless ``index_dtype=torch.int64`` asks for the reference's widening; ``pin`` places the arrays in page-locked
    memory for ``HostRayLoader``; ``device`` uploads them for ``RayDataLoader``; ``mmap`` maps the files instead of
    reading them (ray sets larger than host memory)."""
    def read(stem, dtype):
        arr = np.load(os.path.join(preproc_data_path, stem + ".npy"), mmap_mode="c" if mmap else None)   # "c": private map
        t = torch.from_numpy(arr).to(dtype)
        if pin:
            t = t.pin_memory()
        if device is not None:
            t = t.to(device)
        return t

    data = {"vertex_idxs_of_hit_faces": read("vids_of_hit_faces", index_dtype),
            "barycentric_coords": read("barycentric_coords", torch.float32),
            "expected_rgbs": read("expected_rgbs", torch.float32)}
    dirs_p, face_p = (os.path.join(preproc_data_path, s + ".npy") for s in ("unit_ray_dirs", "face_idxs"))
    if os.path.exists(dirs_p) and os.path.exists(face_p):
        data["unit_ray_dirs"] = read("unit_ray_dirs", torch.float32)
        data["face_idxs"] = read("face_idxs", index_dtype)
    return data


def load_first_k_eigenfunctions(eigenfunctions_path, k, rescale_strategy="standard", embed_strategy=None, eigenvalues_path=None,
                                ts=128) -> torch.Tensor:
    """The table the hot path gathers from, as the reference derives it from the stored eigen-decomposition
    (mesh.py:53-107): column selection (first ``k`` or an explicit list), optional GPS / HKS embedding with the
    eigenvalues, then the rescale (``standard``: each column divided by its range; ``one-norm``: each row by its
    Euclidean norm; ``unscaled``).  The arithmetic runs in the stored dtype (float64 from the eigensolver) and is cast
    to float32 once at the end, like the reference -- the result is bit-identical (tests/test_ray_store.py)."""
    all_efuncs = np.load(eigenfunctions_path)
    if isinstance(k, list):
        efuncs = all_efuncs[:, np.array(k)]
    else:
        if k > all_efuncs.shape[1]:
            raise ValueError(f"k={k} exceeds the {all_efuncs.shape[1]} stored eigenfunctions")
        efuncs = all_efuncs[:, :k]
    evals = None
    if eigenvalues_path is not None:
        all_evals = np.load(eigenvalues_path)
        evals = all_evals[..., np.array(k)] if isinstance(k, list) else all_evals[:k]   # mesh.py:69-75
        if np.abs(evals[0]) < 1e-10 and evals[0] < 0:          # numerically-zero first eigenvalue with the wrong sign
            evals[0] *= -1
        if not np.all(evals > 0):
            raise ValueError(f"eigenvalues must be positive, min value: {evals.min()}")
    if embed_strategy is not None:
        if evals is None:
            raise ValueError(f"embed_strategy={embed_strategy!r} needs eigenvalues_path")
        if embed_strategy == "gps":                            # Global Point Signature
            w = np.sqrt(evals)
            w /= w[0]
            # mesh.py:91 returns here, before any rescale -- as a float64 numpy array, which the reference's loaders cannot
            # move to a device; same values, as the float32 tensor every other branch returns
            return torch.from_numpy(efuncs / w).to(dtype=torch.float32)
        if embed_strategy == "hks":                            # Heat Kernel Signature
            t = np.logspace(-2, 0, num=ts)
            efuncs = (efuncs * efuncs) @ np.exp(-evals[..., None] @ t[None, ...])
        else:
            raise ValueError(f"Unknown embedding strategy {embed_strategy}")
    if rescale_strategy == "standard":
        efuncs = efuncs / (np.max(efuncs, axis=0, keepdims=True) - np.min(efuncs, axis=0, keepdims=True))
    elif rescale_strategy == "one-norm":
        efuncs = efuncs / np.linalg.norm(efuncs, ord=2, axis=-1, keepdims=True)
    elif rescale_strategy != "unscaled":
        raise RuntimeError(f"Unknown rescaling strategy: {rescale_strategy}")
    return torch.from_numpy(efuncs).to(dtype=torch.float32)
