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

from .ops import EigenTable, validate_indices


_RAY_KEYS = ("vertex_idxs_of_hit_faces", "barycentric_coords", "expected_rgbs", "unit_ray_dirs", "hit_face_idxs")


class RayDataLoader:
    """Device-resident split, one batch of gather inputs per ``next()``.

    Constructor arguments and iterator protocol are the reference's (ray_dataloader.py:57-69,100-118).  Per-ray arrays
    live in ``self.arrays`` (key = batch-dict key); a batch is either the index-selected rows of each array (the
    reference's behaviour, ray_dataloader.py:122-126,142-143) or, with ``fused_select``, the arrays themselves plus the
    batch's ray indices, which the kernels apply while loading.
    """

    def __init__(self, features, feature_strategy, vertex_idxs_of_hit_faces, barycentric_coords, expected_rgbs,
                 unit_ray_dirs, face_idxs, batch_size, shuffle, drop_last, device="cuda", table_dtype="f32",
                 materialize=False, generator=None, fused_select=False):
        if feature_strategy != "efuncs":
            raise ValueError(f"Unknown / unsupported input feature strategy for the B200 path: {feature_strategy}")
        if (unit_ray_dirs is None) != (face_idxs is None):
            raise ValueError("unit_ray_dirs and face_idxs come together (view-direction branch) or not at all")
        self.device = torch.device(device)
        self.feature_strategy = feature_strategy
        self.table = features if isinstance(features, EigenTable) else EigenTable.from_dense(
            features.to(torch.float32), device=self.device, dtype=table_dtype)
        given = dict(zip(_RAY_KEYS, (vertex_idxs_of_hit_faces, barycentric_coords, expected_rgbs, unit_ray_dirs, face_idxs)))
        self.arrays = {k: v.to(self.device) for k, v in given.items() if v is not None}      # H2D once, as :73-77
        if self.arrays["barycentric_coords"].dtype != torch.float32:
            raise TypeError("barycentric_coords must be float32 (ray_dataloader.py:132)")
        validate_indices(self.table, self.arrays["vertex_idxs_of_hit_faces"])     # once per split, not per step
        self.shuffle, self.drop_last = bool(shuffle), bool(drop_last)
        self.materialize, self.fused_select, self.generator = materialize, fused_select, generator
        self.B = int(batch_size)
        self.N = int(self.arrays["vertex_idxs_of_hit_faces"].shape[0])
        self.num_batches = self.N // self.B if self.drop_last else -(-self.N // self.B)
        self.order = torch.arange(self.N, device=self.device)
        self.cursor = 0

    # the reference's attribute names, for code that reaches into the loader
    vertex_idxs_of_hit_faces = property(lambda self: self.arrays["vertex_idxs_of_hit_faces"])
    barycentric_coords = property(lambda self: self.arrays["barycentric_coords"])
    expected_rgbs = property(lambda self: self.arrays["expected_rgbs"])
    unit_ray_dirs = property(lambda self: self.arrays.get("unit_ray_dirs"))
    face_idxs = property(lambda self: self.arrays.get("hit_face_idxs"))

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        if self.shuffle:                                       # a fresh permutation per epoch (ray_dataloader.py:105)
            self.order = torch.randperm(self.N, device=self.device, generator=self.generator)
        self.cursor = 0
        return self

    def set_epoch_order(self, idxs: torch.Tensor):
        """Use an explicit ray order for the next epoch (parity runs feed CPU and CUDA the same batches)."""
        self.order, self.cursor = idxs.to(self.device), 0
        return self

    def __next__(self):
        if self.cursor >= self.num_batches:
            raise StopIteration
        sel = self.order[self.cursor * self.B:min((self.cursor + 1) * self.B, self.N)]
        self.cursor += 1
        if self.fused_select:
            batch = dict(self.arrays, ray_idxs=sel)
        else:
            batch = {k: v[sel] for k, v in self.arrays.items()}
        batch["eigenfunction_table"] = self.table
        if self.materialize and not self.fused_select:
            from .ops import gather_interpolate_dense
            batch["eigenfunctions"] = gather_interpolate_dense(self.table, batch["vertex_idxs_of_hit_faces"],
                                                              batch["barycentric_coords"])
        return batch


class HostRayLoader:
    """Batch producer for splits that live in (pinned) HOST memory: the 36-68 B/ray gather inputs are copied to the
    device one batch ahead of the consumer on a dedicated copy stream, so the H2D transfer of batch i+1 overlaps the
    kernels of batch i.  Same batch dict as :class:`RayDataLoader`; same index-select semantics
    (ray_dataloader.py:115-145) when ``shuffle`` is on (the permutation is drawn on the host).

    ``vertex_idxs_of_hit_faces`` / ``face_idxs`` may be int32 exactly as the reference stores them on disk
    (mesh.py:513-516) -- the kernels consume either width and the values pass through bit-exactly -- which halves the
    index bytes on the wire compared with the reference's int64 upcast (dataset.py:16,31).
    """

    def __init__(self, table, vertex_idxs_of_hit_faces, barycentric_coords, expected_rgbs=None, unit_ray_dirs=None,
                 face_idxs=None, batch_size=1 << 17, shuffle=False, drop_last=False, device="cuda", generator=None,
                 depth: int = 2):
        if not isinstance(table, EigenTable):
            raise TypeError("table must be an inf_b200.EigenTable (device resident)")
        self.table = table
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostRayLoader feeds the CUDA path: device must be a CUDA device")

        self.rebind(vertex_idxs_of_hit_faces, barycentric_coords, expected_rgbs, unit_ray_dirs, face_idxs)
        self.N = self.host["vertex_idxs_of_hit_faces"].shape[0]
        self.B = batch_size
        self.shuffle, self.drop_last, self.generator = shuffle, drop_last, generator
        self.num_batches = self.N // self.B if drop_last else (self.N + self.B - 1) // self.B
        if depth < 2:
            raise ValueError("HostRayLoader needs depth >= 2: batch i+1 is copied while batch i is consumed")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.device)
        # device staging ring; each slot: tensors + "copy done" / "consumer done" events
        self.slots = []
        for _ in range(depth):
            bufs = {k: torch.empty((self.B,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device) for k, v in self.host.items()}
            self.slots.append({"bufs": bufs, "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "used": False})
        self.perm = None
        self.h2d_bytes_per_ray = sum(v[0].numel() * v.element_size() for v in self.host.values())
        self._next_issue = self._next_take = 0

    def rebind(self, vertex_idxs_of_hit_faces, barycentric_coords, expected_rgbs=None, unit_ray_dirs=None, face_idxs=None):
        """Point the loader at other host arrays of the same layout (staging buffers, streams and events are kept)."""
        def pin(t):
            if t is None:
                return None
            t = t.contiguous()
            return t if t.is_pinned() else t.pin_memory()

        host = {"vertex_idxs_of_hit_faces": pin(vertex_idxs_of_hit_faces), "barycentric_coords": pin(barycentric_coords),
                "expected_rgbs": pin(expected_rgbs), "unit_ray_dirs": pin(unit_ray_dirs), "hit_face_idxs": pin(face_idxs)}
        host = {k: v for k, v in host.items() if v is not None}
        assert host["barycentric_coords"].dtype == torch.float32
        validate_indices(self.table, host["vertex_idxs_of_hit_faces"])            # once per bound split
        if getattr(self, "slots", None):
            old = self.host
            assert host.keys() == old.keys() and all(host[k].dtype == old[k].dtype and host[k].shape[1:] == old[k].shape[1:]
                                                     for k in host), "rebind needs the same keys / dtypes"
            self.N = host["vertex_idxs_of_hit_faces"].shape[0]
            self.num_batches = self.N // self.B if self.drop_last else (self.N + self.B - 1) // self.B
        self.host = host
        return self

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        self.perm = torch.randperm(self.N, generator=self.generator) if self.shuffle else None
        self._next_issue = self._next_take = 0
        self._pending = False
        # Slots handed out in the previous epoch may still be read by kernels the caller enqueued on the current stream
        # (the last batch of an epoch never gets its "free" event from __next__): every copy of the new epoch waits for
        # everything enqueued so far.
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        for slot in self.slots:
            slot["used"] = False
        for _ in range(min(self.depth, self.num_batches)):
            self._issue()
        return self

    def _issue(self):
        i = self._next_issue
        if i >= self.num_batches:
            return
        lo, hi = i * self.B, min((i + 1) * self.B, self.N)
        slot = self.slots[i % self.depth]
        with torch.cuda.stream(self.copy_stream):
            if slot["used"]:
                self.copy_stream.wait_event(slot["free"])       # the kernels that read this slot have finished
            for k, src in self.host.items():
                h = src[lo:hi] if self.perm is None else src[self.perm[lo:hi]].pin_memory()
                slot["bufs"][k][: hi - lo].copy_(h, non_blocking=True)
            slot["ready"].record(self.copy_stream)
        slot["n"] = hi - lo
        self._next_issue = i + 1

    def __next__(self):
        i = self._next_take
        if i >= self.num_batches:
            raise StopIteration
        self._run_pending()                              # the consumer of the previous batch never called the hook
        slot = self.slots[i % self.depth]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(slot["ready"])
        n = slot["n"]
        batch = {k: (v if n == self.B else v[:n]) for k, v in slot["bufs"].items()}
        batch["eigenfunction_table"] = self.table
        self._next_take = i + 1
        # The slot handed out before this one is reused by the next copy; its consumer is whatever the caller enqueued
        # on the current stream before this call.  Issuing that copy costs ~50 us of host time, so it is deferred: the
        # trainer steps call batch["on_launched"]() right after they have enqueued this batch's kernels (the host
        # then works while the GPU does); if nobody calls it, the next __next__ does.
        if i >= 1:
            prev = self.slots[(i - 1) % self.depth]
            prev["free"].record(cur)
            prev["used"] = True
            self._pending = True
            batch["on_launched"] = self._run_pending
        return batch

    def _run_pending(self):
        if getattr(self, "_pending", False):
            self._pending = False
            self._issue()


def create_ray_dataloader(preproc_data_path, eigenfunctions_path, k, feature_strategy, mesh, rescale_strategy,
                          eigenvalues_path, embed_strategy, batch_size, shuffle, drop_last, device="cuda",
                          table=None, fused_select=False, group=None):
    """ray_dataloader.py:7-58: eigenfunction table + preprocessed-ray directory -> ``RayDataLoader``.  Positional
    order as the reference declares it (``eigenvalues_path`` before ``embed_strategy``).  The indices stay int32 as
    stored (the kernels read either width); ``table`` lets several splits share one device-resident table instead of
    uploading it once per loader as the reference does (ray_dataloader.py:73)."""
    from .ray_store import load_first_k_eigenfunctions, load_preprocessed_rays
    if feature_strategy != "efuncs":
        raise ValueError(f"Unknown / unsupported input feature strategy for the B200 path: {feature_strategy}")
    if table is None:
        table = load_first_k_eigenfunctions(eigenfunctions_path, k, rescale_strategy=rescale_strategy,
                                            embed_strategy=embed_strategy, eigenvalues_path=eigenvalues_path)
    d = load_preprocessed_rays(preproc_data_path)
    from . import parallel
    world = parallel.world_size(group)
    if world > 1:      # this rank's contiguous 1/W of the split and of every batch
        n_all = d["expected_rgbs"].shape[0]
        if shuffle:                                        # training split: equal shards, every rank runs the same number of steps
            each = n_all // world
            lo, hi = parallel.rank(group) * each, parallel.rank(group) * each + each
        else:                                              # val / test: every ray is evaluated exactly once (ragged tail kept);
            lo, hi = parallel.shard_range(n_all, group)    # the epoch metrics weight by the all-reduced ray count
        d = {k_: v[lo:hi] for k_, v in d.items()}
        batch_size = max(1, batch_size // world)
    return RayDataLoader(table, feature_strategy, d["vertex_idxs_of_hit_faces"], d["barycentric_coords"], d["expected_rgbs"],
                         d.get("unit_ray_dirs"), d.get("face_idxs"), batch_size, shuffle, drop_last, device=device,
                         fused_select=fused_select)


def get_data(config, device="cuda", fused_select=True, group=None):
    """config.py:58-103: the train / val (/ test) loaders of a config.  The table is built and uploaded once and shared
    by the splits.  (The reference passes ``embed_strategy`` and ``eigenvalues_path`` to ``create_ray_dataloader`` in
    swapped positions, config.py:66-67 vs ray_dataloader.py:13-14; they are passed by name here.)"""
    from .ray_store import load_first_k_eigenfunctions
    dc, mc = config["data"], config["model"]
    strategy = mc.get("feature_strategy", "efuncs")
    if strategy != "efuncs":
        raise ValueError(f"Unknown / unsupported input feature strategy for the B200 path: {strategy}")
    feats = load_first_k_eigenfunctions(dc["eigenfunctions_path"], mc.get("k"), rescale_strategy=dc.get("rescale_strategy", "standard"),
                                        embed_strategy=dc.get("embed_strategy"), eigenvalues_path=dc.get("eigenvalues_path"))
    table = EigenTable.from_dense(feats, device=torch.device(device))
    B = config["training"]["batch_size"]

    def mk(path, shuffle, drop_last, fs):
        return create_ray_dataloader(path, None, None, strategy, None, None, None, None, B, shuffle, drop_last, device=device,
                                     table=table, fused_select=fs, group=group)
    data = {"train": mk(dc["preproc_data_path_train"], True, dc.get("train_drop_last", True), fused_select),
            "val": mk(dc["preproc_data_path_eval"], False, False, False)}
    if dc.get("preproc_data_path_test") is not None:
        data["test"] = mk(dc["preproc_data_path_test"], False, False, False)
    return data
