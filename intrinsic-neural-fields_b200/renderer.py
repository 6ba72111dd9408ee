"""Inference data path of the reference's ``Renderer.render`` (renderer.py:64-146), minus ray casting.

Ray generation and ray-mesh intersection stay the reference's CPU preprocessing (trimesh/pyembree,
mesh.py:171-310); they produce, per hit, the vertex-id triplet, the barycentric weights, the hit-ray
index (and optionally ray direction / face id).  ``FieldRenderer.render_hits`` is everything after that:
instead of gathering on the CPU and uploading a 4 KB/ray embedding (mesh.py:327-339, renderer.py:116) it
uploads the 36 B/ray gather inputs, runs the fused kernel in chunks against the resident table, and
scatters the RGB into the canvas by ``hit_ray_idxs`` exactly as renderer.py:119-142 does.
"""
from __future__ import annotations

from typing import Optional

import torch

from .ops import EigenTable, validate_indices


class GraphedForward:
    """``model(batch)`` under ``torch.no_grad()`` replayed from a CUDA graph, one graph per distinct set of input buffers.

    The kernels of an inference call take ~0.1 ms at 131 072 rays, the eager Python path into them about twice that, so a
    renderer that streams frames through fixed staging buffers (``FieldRenderer.predict_hits_host``) replays instead.
    The first call with a set of buffers runs eagerly (sizes the workspace), the second captures, later ones replay.
    The returned tensor is the graph's static output: consume (or copy) it before the same buffers are used again --
    ``mark_consumed`` lets a consumer on another stream tell the next replay when it has finished reading."""

    def __init__(self, model):
        self.model, self.graphs = model, {}

    def _key(self, batch):
        # a re-pack of the weights (new operand images, new vertex projection) must not replay a graph of the old buffers
        return tuple((k, v.data_ptr(), tuple(v.shape)) for k, v in sorted(batch.items()) if torch.is_tensor(v)) + \
            (id(batch.get("eigenfunction_table")), getattr(self.model, "_pack_gen", 0))

    @torch.no_grad()
    def __call__(self, batch):
        self.model.eval()
        self.model.prepare(batch)          # operand images / vertex projection of the current weights (cached)
        key = self._key(batch)
        ent = self.graphs.get(key)
        if ent is None:
            self.graphs[key] = "warm"
            return self.model(batch)
        if ent == "warm":
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.model(batch)
            ent = self.graphs[key] = {"graph": g, "out": out, "done_event": None}
        if ent["done_event"] is not None:
            torch.cuda.current_stream().wait_event(ent["done_event"])      # the previous output of this graph has been read
        ent["graph"].replay()
        return ent["out"]

    def mark_consumed(self, batch, event):
        ent = self.graphs.get(self._key(batch))
        if isinstance(ent, dict):
            ent["done_event"] = event


class FieldRenderer:
    def __init__(self, model, eigenfunctions, H: int, W: int, background: str = "white", device="cuda",
                 chunk: int = 1 << 20, table_dtype: str = "f32"):
        assert background in ("white", "black")
        self.model, self.H, self.W, self.background = model, H, W, background
        self.device = torch.device(device)
        self.chunk = chunk
        self.set_eigenfunctions(eigenfunctions, table_dtype)

    def set_eigenfunctions(self, eigenfunctions, table_dtype: str = "f32"):
        """Swap the table under a fixed model (texture transfer passes a mapped table per renderer)."""
        self.table = eigenfunctions if isinstance(eigenfunctions, EigenTable) else EigenTable.from_dense(
            eigenfunctions.to(torch.float32), device=self.device, dtype=table_dtype)

    def _validate(self, vids, face_idxs):
        """Range-check a frame's index arrays once, where it enters (the kernels dereference them raw)."""
        normals = getattr(self.model, "face_normals", None)
        validate_indices(self.table, vids, face_idxs, None if normals is None else int(normals.shape[0]))

    @torch.no_grad()
    def predict_hits(self, vertex_idxs_of_hit_faces: torch.Tensor, barycentric_coords: torch.Tensor,
                     unit_ray_dirs: Optional[torch.Tensor] = None, face_idxs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(M,3) vids + (M,3) bary on the host or device -> (M,3) RGB on the device, chunked."""
        self.model.eval()
        M = vertex_idxs_of_hit_faces.shape[0]
        self._validate(vertex_idxs_of_hit_faces, face_idxs)
        out = torch.empty((M, 3), dtype=torch.float32, device=self.device)
        for lo in range(0, M, self.chunk):
            hi = min(M, lo + self.chunk)
            batch = {"eigenfunction_table": self.table,
                     "vertex_idxs_of_hit_faces": vertex_idxs_of_hit_faces[lo:hi].to(self.device, non_blocking=True),
                     "barycentric_coords": barycentric_coords[lo:hi].to(self.device, non_blocking=True)}
            if unit_ray_dirs is not None:
                batch["unit_ray_dirs"] = unit_ray_dirs[lo:hi].to(self.device, non_blocking=True)
                batch["hit_face_idxs"] = face_idxs[lo:hi].to(self.device, non_blocking=True)
            out[lo:hi] = self.model(batch)
        return out

    @torch.no_grad()
    def predict_hits_host(self, vertex_idxs_of_hit_faces: torch.Tensor, barycentric_coords: torch.Tensor,
                          unit_ray_dirs: Optional[torch.Tensor] = None, face_idxs: Optional[torch.Tensor] = None,
                          chunk: int = 1 << 18, out: Optional[torch.Tensor] = None, use_graphs: bool = True) -> torch.Tensor:
        """Host -> host inference of renderer.py:113-119: gather inputs in (pinned) host memory, RGB back in pinned
        host memory.  The reference uploads the 4 KB/ray embedding per 2^15-ray chunk and blocks on ``.cpu()``; here the
        36-68 B/ray inputs of chunk i+1 are copied while chunk i computes and the RGB of chunk i-1 is copied back, on
        dedicated copy streams.  Returns after the last D2H copy has completed.  Chunks of 2^18 rays keep the host's
        per-chunk launch path (~0.1 ms) off the critical path: measured host to host, 0.93 G rays/s at 2^17-ray chunks,
        1.50 G at 2^18, 1.57 G at 2^19 = 56 GB/s over PCIe, the link's limit (scripts/infer_e2e_chunks.py)."""
        from .ray_dataloader import HostRayLoader
        self.model.eval()
        M = vertex_idxs_of_hit_faces.shape[0]
        if out is None:
            out = torch.empty((M, 3), dtype=torch.float32).pin_memory()
        if M == 0:
            return out
        self._validate(None, face_idxs)                   # the vertex ids are checked where the loader binds them
        key = (min(chunk, M), vertex_idxs_of_hit_faces.dtype, unit_ray_dirs is not None, face_idxs is not None)
        if getattr(self, "_hl_key", None) == key and self._hl.table is self.table:
            loader = self._hl.rebind(vertex_idxs_of_hit_faces, barycentric_coords, None, unit_ray_dirs, face_idxs)
        else:
            loader = HostRayLoader(self.table, vertex_idxs_of_hit_faces, barycentric_coords, None, unit_ray_dirs, face_idxs,
                                   batch_size=min(chunk, M), device=self.device, depth=3)
            self._hl, self._hl_key = loader, key
        if not hasattr(self, "_d2h"):
            self._d2h = torch.cuda.Stream(self.device)
        cur = torch.cuda.current_stream(self.device)
        lo = 0
        keep = []
        if not hasattr(self, "_gf") or self._gf.model is not self.model:
            self._gf = GraphedForward(self.model)
        for batch in loader:
            hook = batch.pop("on_launched", None)
            full = batch["barycentric_coords"].shape[0] == loader.B
            # full chunks live in the loader's staging slots: one CUDA-graph launch per chunk (ragged tail: eager)
            pred = self._gf(batch) if (use_graphs and full) else self.model(batch)
            if hook is not None:
                hook()                                   # kernels enqueued: let the loader issue its next H2D copies
            n = pred.shape[0]
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(self._d2h):
                self._d2h.wait_event(ev)
                out[lo:lo + n].copy_(pred, non_blocking=True)
                if use_graphs and full:
                    read = torch.cuda.Event()
                    read.record(self._d2h)
                    self._gf.mark_consumed(batch, read)  # the slot's next replay waits until this RGB has left the device
            pred.record_stream(self._d2h)
            keep.append(pred)
            lo += n
        self._d2h.synchronize()
        return out

    @torch.no_grad()
    def canvas_on_device(self, vertex_idxs_of_hit_faces, barycentric_coords, hit_ray_idxs, obj_mask_1d=None,
                         unit_ray_dirs=None, face_idxs=None) -> torch.Tensor:
        """Canvas assembly of renderer.py:119-146 on the device: (H,W,3) float32 CUDA tensor, everything enqueued
        asynchronously on the current stream (no host synchronisation)."""
        assert obj_mask_1d is None or obj_mask_1d.size()[0] == self.H * self.W
        pred = self.predict_hits(vertex_idxs_of_hit_faces, barycentric_coords, unit_ray_dirs, face_idxs)
        fill = 1.0 if self.background == "white" else 0.0
        N = self.H * self.W if obj_mask_1d is None else int(obj_mask_1d.sum())
        img = torch.full((N, 3), fill, dtype=torch.float32, device=self.device)
        img[hit_ray_idxs.to(self.device)] = pred
        if obj_mask_1d is not None:
            assert obj_mask_1d.dtype == torch.bool
            full = torch.full((self.H * self.W, 3), fill, dtype=torch.float32, device=self.device)
            full[obj_mask_1d.to(self.device)] = img
            img = full
        return img.reshape(self.H, self.W, 3)

    @torch.no_grad()
    def render_hits(self, vertex_idxs_of_hit_faces, barycentric_coords, hit_ray_idxs, obj_mask_1d=None,
                    unit_ray_dirs=None, face_idxs=None, eval_render: bool = False):
        """Canvas assembly of renderer.py:119-146 on the device; returns (H,W,3) numpy (or torch + hit idxs)."""
        img = self.canvas_on_device(vertex_idxs_of_hit_faces, barycentric_coords, hit_ray_idxs, obj_mask_1d,
                                    unit_ray_dirs, face_idxs).cpu()
        if eval_render:
            return img, hit_ray_idxs
        return img.numpy()


class Renderer(FieldRenderer):
    """The reference's ``Renderer`` call surface (renderer.py:36-146): same constructor arguments, ``set_height`` /
    ``set_width`` / ``apply_mesh_transform`` and ``render(camCv2world, K, obj_mask_1d, eval_render, ...)``.

    Ray casting is not part of this package: ``intersect(ray_origins, unit_ray_dirs)`` is the caller's
    ``ray_mesh_intersect`` (mesh.py:208-245; trimesh + pyembree in the reference) returning ``(vids (M,3), bary (M,3),
    hit_ray_idxs (M,), face_idxs (M,))``.  Everything around it -- pixel rays (mesh.py:171-205), the 36-68 B/ray upload
    instead of the CPU gather + 4 KB/ray embedding (mesh.py:327-339, renderer.py:116), chunked inference, the canvas
    scatter (renderer.py:119-146) -- runs here.
    """

    def __init__(self, model, mesh, eigenfunctions=None, feature_strategy="efuncs", background="white", device="cuda",
                 *, H, W, intersect=None, table_dtype: str = "f32"):
        if feature_strategy != "efuncs":
            raise NotImplementedError(f"feature strategy {feature_strategy!r}: only the intrinsic 'efuncs' path is built")
        if eigenfunctions is None:
            raise ValueError("the 'efuncs' strategy needs the eigenfunction table")
        super().__init__(model, eigenfunctions, H, W, background=background, device=device, table_dtype=table_dtype)
        self.mesh, self.feature_strategy, self.intersect = mesh, feature_strategy, intersect

    def set_height(self, height):
        self.H = height

    def set_width(self, width):
        self.W = width

    def apply_mesh_transform(self, transform):
        self.mesh.apply_transform(transform)
        if hasattr(self.intersect, "rebuild"):          # the reference rebuilds its intersector here (renderer.py:62)
            self.intersect.rebuild(self.mesh)

    def cast(self, camCv2world, K, obj_mask_1d=None, distortion_coeffs=None, distortion_type=None, pixel_coords=None):
        """Host side of a frame (renderer.py:83-112): pixel rays of the masked pixels and the caller's intersector.
        Returns the arguments of ``canvas_on_device``."""
        from .ray_store import pixel_rays
        if self.intersect is None:
            raise RuntimeError("Renderer.render needs an intersector callable (ray casting is outside this package)")
        if distortion_type is not None and pixel_coords is None:
            raise NotImplementedError("distorted views: undistort the masked pixels first and pass them as pixel_coords")
        assert obj_mask_1d is None or obj_mask_1d.size()[0] == self.H * self.W
        mask = torch.ones(self.H * self.W, dtype=torch.bool) if obj_mask_1d is None else obj_mask_1d
        origins, dirs = pixel_rays(camCv2world, K, mask, H=self.H, W=self.W, pixel_coords=pixel_coords)
        vids, bary, hit_ray_idxs, face_idxs = self.intersect(origins, dirs)
        assert vids.shape[0] > 0                                                         # renderer.py:110
        view = getattr(self.model, "cfg", None) is not None and self.model.cfg.has_view
        return dict(vertex_idxs_of_hit_faces=vids, barycentric_coords=bary.to(torch.float32), hit_ray_idxs=hit_ray_idxs,
                    obj_mask_1d=obj_mask_1d, unit_ray_dirs=dirs[hit_ray_idxs] if view else None,
                    face_idxs=face_idxs if view else None)

    @torch.no_grad()
    def render(self, camCv2world, K, obj_mask_1d=None, eval_render=False, distortion_coeffs=None, distortion_type=None,
               pixel_coords=None):
        hits = self.cast(camCv2world, K, obj_mask_1d, distortion_coeffs, distortion_type, pixel_coords)
        return self.render_hits(eval_render=eval_render, **hits)

    @torch.no_grad()
    def render_frames(self, views):
        """The frame loop of render_video.py:117-134 as a pipeline: yields one (H,W,3) numpy image per view, in order.

        ``views`` yields ``(camCv2world, K)`` or ``(camCv2world, K, obj_mask_1d)``.  While the GPU evaluates frame n (upload
        of the 36-68 B/ray gather inputs, the fused kernels, the canvas scatter -- all on the device) a host thread already
        runs the caller's intersector for frame n+1, and frame n-1's canvas travels to pinned host memory on a copy
        stream; the only host synchronisation per frame is the wait for that copy.  Every frame goes through the same
        ``canvas_on_device`` as ``render()``, so the images are bit-identical to calling ``render()`` view by view."""
        from concurrent.futures import ThreadPoolExecutor
        views = iter(views)

        def cast_next():
            v = next(views, None)
            return None if v is None else self.cast(v[0], v[1], v[2] if len(v) > 2 else None)

        d2h = torch.cuda.Stream(self.device)
        bufs = [torch.empty((self.H, self.W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        pending = None                                   # (buffer, event) of the frame whose canvas is still on its way
        with ThreadPoolExecutor(max_workers=1) as pool:
            fut = pool.submit(cast_next)
            n = 0
            while True:
                hits = fut.result()
                if hits is None:
                    break
                fut = pool.submit(cast_next)             # frame n+1 is cast on the host while the GPU works on frame n
                canvas = self.canvas_on_device(**hits)
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.device))
                buf, ev = bufs[n % 2], torch.cuda.Event()
                with torch.cuda.stream(d2h):
                    d2h.wait_event(done)
                    buf.copy_(canvas, non_blocking=True)
                    ev.record(d2h)
                canvas.record_stream(d2h)
                if pending is not None:
                    pending[1].synchronize()             # the one host wait per frame
                    yield pending[0].numpy().copy()
                pending = (buf, ev)
                n += 1
            if pending is not None:
                pending[1].synchronize()
                yield pending[0].numpy().copy()
