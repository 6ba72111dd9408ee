"""Texel prediction for UV baking / texture transfer through the fused path.

``bake_texture_field.py:266-319`` evaluates the trained field at every texel's surface point: a numpy gather of
the three eigenfunction rows, a float64 ``einsum`` with the barycentric weights, a float32 upload of the (n,k)
embedding per 2^15-texel chunk, ``model(batch)`` and a blocking ``.cpu()``.  Here the (face-vertex ids, barycentric
weights) pairs go to the device as they are (36 B/texel) and the fused kernel gathers from the resident table.
"""
from __future__ import annotations

import numpy as np
import torch

from .ops import EigenTable
from .renderer import FieldRenderer


@torch.no_grad()
def pred_rgbs(model, features, faces_index_efs, barycentric_coords, device="cuda", chunk: int = 1 << 17,
              table_dtype: str = "f32") -> np.ndarray:
    """(n,3) integer vertex ids + (n,3) barycentric weights (numpy, any float width) -> (n,3) f32 RGB (numpy).

    ``features`` is the (V,k) eigenfunction table (numpy / torch / an ``EigenTable`` already on the device).
    View dependence is rejected like the reference does (bake_texture_field.py:271-273): a texel has no view ray.
    """
    if getattr(model, "cfg", None) is not None and model.cfg.has_view:
        raise NotImplementedError("Currently view dependence is not supported.")
    if not isinstance(features, EigenTable):
        features = torch.as_tensor(features, dtype=torch.float32)
    vids = np.ascontiguousarray(faces_index_efs)
    if vids.dtype not in (np.int32, np.int64):
        vids = vids.astype(np.int64)
    bary = np.ascontiguousarray(barycentric_coords, dtype=np.float32)
    r = FieldRenderer(model, features, H=1, W=1, device=device, table_dtype=table_dtype)
    out = r.predict_hits_host(torch.from_numpy(vids).pin_memory(), torch.from_numpy(bary).pin_memory(), chunk=chunk)
    return out.numpy().copy()
