"""Batch producer mirroring the reference's ``RayDataLoader`` (ray_dataloader.py:57-145).

Same constructor arguments and iterator protocol; the whole split lives on the device and each batch is
selected with the same index tensors.  The difference is what a batch carries: instead of the
materialised ``"eigenfunctions"`` (B,k) embedding it carries the *inputs* of the gather --
``"vertex_idxs_of_hit_faces"`` (B,3) and ``"barycentric_coords"`` (B,3), passed through bit-exactly --
plus a handle to the device-resident padded table, so the fused kernel interpolates on the fly.
``materialize=True`` additionally emits the reference's ``"eigenfunctions"`` key (dense path).
"""
from __future__ import annotations

import torch

from .ops import EigenTable


class RayDataLoader:
    def __init__(self, features, feature_strategy, vertex_idxs_of_hit_faces, barycentric_coords, expected_rgbs,
                 unit_ray_dirs, face_idxs, batch_size, shuffle, drop_last, device="cuda", table_dtype="f32",
                 materialize=False, generator=None):
        if feature_strategy != "efuncs":
            raise ValueError(f"Unknown / unsupported input feature strategy for the B200 path: {feature_strategy}")
        self.device = torch.device(device)
        self.feature_strategy = feature_strategy
        self.table = features if isinstance(features, EigenTable) else EigenTable.from_dense(
            features.to(torch.float32), device=self.device, dtype=table_dtype)
        self.vertex_idxs_of_hit_faces = vertex_idxs_of_hit_faces.to(self.device)
        self.barycentric_coords = barycentric_coords.to(self.device)
        self.expected_rgbs = expected_rgbs.to(self.device)
        self.unit_ray_dirs = unit_ray_dirs
        self.face_idxs = face_idxs
        if self.unit_ray_dirs is not None:
            assert self.face_idxs is not None
            self.unit_ray_dirs = self.unit_ray_dirs.to(self.device)
            self.face_idxs = self.face_idxs.to(self.device)
        assert self.barycentric_coords.dtype == torch.float32
        self.shuffle, self.drop_last = shuffle, drop_last
        self.materialize = materialize
        self.generator = generator
        self.B = batch_size
        self.N = self.vertex_idxs_of_hit_faces.shape[0]
        self.num_batches = self.N // self.B if drop_last else (self.N + self.B - 1) // self.B
        self.i = 0
        self.idxs = torch.arange(self.N, device=self.device)

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        if self.shuffle:
            self.idxs = torch.randperm(self.N, device=self.device, generator=self.generator)
        self.i = 0
        return self

    def set_epoch_order(self, idxs: torch.Tensor):
        """Use an explicit ray order for the next epoch (parity runs feed CPU and CUDA the same batches)."""
        self.idxs = idxs.to(self.device)
        self.i = 0
        return self

    def _get_next_batch_idxs(self):
        low = self.i * self.B
        high = min((self.i + 1) * self.B, self.N)
        self.i += 1
        return self.idxs[low:high]

    def __next__(self):
        if self.i >= self.num_batches:
            raise StopIteration
        batch_idxs = self._get_next_batch_idxs()
        batch = {
            "expected_rgbs": self.expected_rgbs[batch_idxs],
            "vertex_idxs_of_hit_faces": self.vertex_idxs_of_hit_faces[batch_idxs],
            "barycentric_coords": self.barycentric_coords[batch_idxs],
            "eigenfunction_table": self.table,
        }
        if self.materialize:
            from .ops import gather_interpolate_dense
            batch["eigenfunctions"] = gather_interpolate_dense(self.table, batch["vertex_idxs_of_hit_faces"],
                                                              batch["barycentric_coords"])
        if self.unit_ray_dirs is not None:
            batch["unit_ray_dirs"] = self.unit_ray_dirs[batch_idxs]
            batch["hit_face_idxs"] = self.face_idxs[batch_idxs]
        return batch
