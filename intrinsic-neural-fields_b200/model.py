"""Drop-in model surface of the reference's ``model.py`` for the intrinsic (eigenfunction) path.

``make_model(model_config, mesh=None)`` returns a module with the same constructor config, the same
``state_dict`` keys/shapes (``layers.0.0.weight``, ``layers.<skip>.Lx.weight``, ``layers.<skip>.Ly.weight``
..., model.py:42-96 / layers.py:50-58 of the reference) and the same ``forward(batch) -> (B,3)``
contract (model.py:98-112), but the arithmetic runs in the fused sm_100a kernels:

* fused path  -- the batch carries ``"vertex_idxs_of_hit_faces"`` (B,3), ``"barycentric_coords"`` (B,3)
  and ``"eigenfunction_table"`` (an :class:`EigenTable`); the interpolated embedding is never built;
* dense path  -- the batch carries the reference's ``"eigenfunctions"`` (B,k) tensor; the same kernels
  read it as the A operand (it is padded to a 16-byte aligned row stride first).

Training: ``forward`` is differentiable w.r.t. the parameters (the backward pass re-runs the fused
forward+backward kernel with the upstream gradient), and :meth:`TextureField.l1_step` is the fast path
used by the trainer for ``loss_type: L1`` (forward + L1 loss + backward in one fused pass,
trainer.py:71-84).  Unsupported reference options (batchnorm, sine, ff/rff/xyz features) raise.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .ops import EigenTable, FieldConfig

RGB_COLOR_DIM = 3  # model.py:9


class _SkipLinear(nn.Module):
    """Parameter container with the reference's key names (layers.py:50-58)."""

    def __init__(self, x_in_dim: int, y_in_dim: int, out_dim: int):
        super().__init__()
        self.Lx = nn.Linear(x_in_dim, out_dim)
        self.Ly = nn.Linear(y_in_dim, out_dim)


def _linear_block(in_dim: int, out_dim: int) -> nn.Sequential:
    # index 0 holds the Linear exactly as in the reference's nn.Sequential(Linear, activation)
    return nn.Sequential(nn.Linear(in_dim, out_dim))


class _FieldFunction(torch.autograd.Function):
    """pred = field(table, vids, bary[, dirs, face_idxs]; params).  No gradient flows into table / vids / bary /
    dirs (none requires grad in the reference either)."""

    @staticmethod
    def forward(ctx, module, table, vids, bary, dense, dirs, face_idxs, face_normals, ray_idxs, extras, *params):
        packed = module._packed_weights(force=True, col_exp=extras[2])
        ctx.module, ctx.dense, ctx.extras = module, dense, extras
        ctx.save_for_backward(table, vids, bary, dirs, face_idxs, face_normals, ray_idxs)
        return torch.ops.inf_b200.field_forward(table, vids, bary, packed, dirs, face_idxs, face_normals,
                                                module.cfg.key(), dense, ray_idxs, *extras)

    @staticmethod
    def backward(ctx, grad_pred):
        module = ctx.module
        table, vids, bary, dirs, face_idxs, face_normals, ray_idxs = ctx.saved_tensors
        _, flat_full = torch.ops.inf_b200.field_forward_backward(
            table, vids, bary, module._packed_weights(col_exp=ctx.extras[2]), None, grad_pred.contiguous().float(), dirs, face_idxs,
            face_normals, module.cfg.key(), ctx.dense, 1.0, ray_idxs, *ctx.extras)
        flat, _ = ops.split_flat(flat_full)
        grads = [flat[off:off + r * c].view(p.shape) for (off, r, c), p in zip(module._layout, module._ordered_params())]
        return (None,) * 10 + tuple(grads)


class _FusedFieldBase(nn.Module):
    """Shared host logic of the two drop-in modules: operand packing cache, batch decoding, forward through the
    custom ops, and the fused L1 training step."""

    cfg: FieldConfig

    def _init_cache(self):
        self._layout = None
        self._packed = None
        self._packed_key = None

    def _ordered_params(self) -> List[torch.Tensor]:
        raise NotImplementedError

    def _view_inputs(self, batch):
        return None, None, None

    def _packed_weights(self, force: bool = False, col_exp: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Operand images of the current parameter values.

        Re-packed (one 5 us kernel) on every training-mode call: in-place optimiser kernels such as
        ``torch.optim.Adam(fused=True)`` update the parameters WITHOUT bumping their version counters, so a
        version-keyed cache would silently serve stale weights.  In eval mode the images are cached, keyed on the
        parameters' storage and version; ``train()`` / ``eval()`` and :meth:`invalidate_packed` drop the cache."""
        params = self._ordered_params()
        # the images are specific to the table's column exponents (EigenTable.col_exp: fp16 range guard)
        key = tuple((p.data_ptr(), p._version) for p in params) + (None if col_exp is None else col_exp.data_ptr(),)
        if force or self._packed is None or self._packed_key != key or self._packed.device != params[0].device:
            with torch.no_grad():
                self._packed = ops.pack_weights(self.cfg, [p.detach().contiguous() for p in params], None, col_exp)
            self._packed_key = key
            self._pack_gen = getattr(self, "_pack_gen", 0) + 1   # every re-pack invalidates what was derived from the images
        if self._layout is None:
            self._layout = ops.param_layout(self.cfg)
        return self._packed

    def invalidate_packed(self):
        self._packed_key = None
        self._proj_key = None

    def _projection(self, batch, packed):
        """Eval mode: the vertex projection P = E [W0;Wy]^T of the cached operand images, computed once per (weights, table)
        and handed to every inference call -- a frozen model renders many chunks / frames (renderer.py:36-146,
        render_video.py:117-134) and then launches the per-ray kernel only.  None when the call cannot use it."""
        table = batch.get("eigenfunction_table")
        if self.training or not isinstance(table, EigenTable) or table.image is None or table.path == "ray" \
                or "vertex_idxs_of_hit_faces" not in batch:
            return None
        key = (getattr(self, "_pack_gen", 0), id(table), table.data.data_ptr())
        if getattr(self, "_proj", None) is None or getattr(self, "_proj_key", None) != key:
            old = getattr(self, "_proj", None)
            if old is not None and (old.device != table.device or old.numel() != ops.projection_bytes(table.n_rows)):
                old = None
            self._proj = ops.project_vertices_raw(self.cfg, table, packed, out=old)   # in place when possible: same address
            self._proj_key = key
        return self._proj

    def train(self, mode: bool = True):
        if mode != self.training:          # a mode switch drops the cached operand images (see _packed_weights)
            self._packed_key = None
        return super().train(mode)

    def prepare(self, batch):
        """(operand images, vertex projection or None) an inference call on ``batch`` would use, brought up to date."""
        extras = self._table_extras(batch)
        packed = self._packed_weights(force=self.training, col_exp=extras[2])
        dense = "eigenfunctions" in batch and "eigenfunction_table" not in batch
        return packed, (None if dense else self._projection(batch, packed))

    def __deepcopy__(self, memo):
        # trainer.py:67,274 deep-copies the model; cached device buffers are rebuilt lazily
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, None if k in ("_packed", "_packed_key", "_proj", "_proj_key") else copy.deepcopy(v, memo))
        return new

    # ---- batch decoding ----------------------------------------------------------------------------
    def _inputs(self, batch: Dict[str, torch.Tensor]):
        table = batch.get("eigenfunction_table")
        if table is not None and "vertex_idxs_of_hit_faces" in batch:
            if not isinstance(table, EigenTable):
                raise TypeError("batch['eigenfunction_table'] must be an inf_b200.EigenTable")
            if table.k != self.cfg.k:
                raise ValueError(f"table has k={table.k}, model expects {self.cfg.k}")
            return table.data, batch["vertex_idxs_of_hit_faces"].contiguous(), batch["barycentric_coords"].contiguous(), False
        if "eigenfunctions" in batch:
            x = batch["eigenfunctions"]
            if x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != self.cfg.k:
                raise ValueError(f"batch['eigenfunctions'] must be (B,{self.cfg.k}) float32")
            return EigenTable.from_dense(x, device=x.device, dtype="f32", image=False, range_guard=False).data, None, None, True
        raise KeyError("batch needs 'eigenfunction_table' + 'vertex_idxs_of_hit_faces' + 'barycentric_coords' "
                       "(fused path) or 'eigenfunctions' (dense path)")

    @staticmethod
    def _table_extras(batch):
        """(fp16 operand image of the table or None, InfDesc.path code, column exponents or None)."""
        table = batch.get("eigenfunction_table")
        if isinstance(table, EigenTable) and "vertex_idxs_of_hit_faces" in batch:
            return table.image, ops.PATHS[table.path], table.col_exp
        return None, 0, None

    @staticmethod
    def _ray_idxs(batch):
        """Optional batch selection folded into the kernels: the per-ray tensors of the batch are then the whole
        split's arrays and ``batch['ray_idxs']`` (B,) picks this batch's rows (ray_dataloader.py:103-145)."""
        r = batch.get("ray_idxs")
        return None if r is None else r.contiguous()

    def forward(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        table, vids, bary, dense = self._inputs(batch)
        dirs, face_idxs, normals = self._view_inputs(batch)
        ridx = self._ray_idxs(batch)
        extras = self._table_extras(batch)
        params = self._ordered_params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _FieldFunction.apply(self, table, vids, bary, dense, dirs, face_idxs, normals, ridx, extras, *params)
        packed = self._packed_weights(force=self.training, col_exp=extras[2])
        proj = None if dense else self._projection(batch, packed)
        return torch.ops.inf_b200.field_forward(table, vids, bary, packed, dirs, face_idxs, normals, self.cfg.key(), dense, ridx,
                                                *extras, proj)

    # ---- fused training pass -----------------------------------------------------------------------
    def l1_step(self, batch: Dict[str, torch.Tensor], global_rays: Optional[int] = None,
                flat_grad: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Forward + mean-L1 loss + backward in one fused pass (trainer.py:75-81, config.py:117-118).

        Sets ``.grad`` of every parameter (views into one flat buffer, ready for a single allreduce) and
        returns ``(loss_sums, pred, flat_grad)`` where ``loss_sums = [sum|pred-target|, sum (pred-target)^2, mean L1 loss]``
        over this call's rays; the mean L1 loss is ``loss_sums[0] / (3 * global_rays)``.
        ``global_rays`` is the number of rays of the whole (all-rank) batch; it defaults to this call's B.
        """
        table, vids, bary, dense = self._inputs(batch)
        dirs, face_idxs, normals = self._view_inputs(batch)
        target = batch["expected_rgbs"].contiguous()
        ridx = self._ray_idxs(batch)
        B = target.shape[0] if ridx is None else ridx.shape[0]
        scale = 1.0 / (3.0 * float(global_rays if global_rays is not None else B))
        extras = self._table_extras(batch)
        pred, flat_full = torch.ops.inf_b200.field_forward_backward(
            table, vids, bary, self._packed_weights(force=True, col_exp=extras[2]), target, None, dirs, face_idxs, normals,
            self.cfg.key(), dense, scale, ridx, *extras)
        flat, sums = ops.split_flat(flat_full)
        if flat_grad is not None:
            flat_grad.copy_(flat)
            flat = flat_grad
        for (off, r, c), p in zip(self._layout, self._ordered_params()):
            p.grad = flat[off:off + r * c].view(p.shape)
        return sums, pred, flat


def _check_common(hidden_dim, input_feature_embed, batchnorm, activation):
    if input_feature_embed not in (None, "efuncs"):
        raise NotImplementedError(f"feature_strategy={input_feature_embed!r}: only the intrinsic 'efuncs' path is "
                                  "implemented by the B200 kernels (extrinsic ff/rff/xyz baselines are out of scope)")
    if batchnorm:
        raise NotImplementedError("batchnorm=True is not supported by the B200 kernels (no shipped config uses it)")
    if activation is not nn.ReLU:
        raise NotImplementedError("only activation='relu' is supported by the B200 kernels")
    if hidden_dim != ops.HIDDEN:
        raise NotImplementedError(f"mlp_hidden_dim={hidden_dim}: compiled set is {{{ops.HIDDEN}}}")


def _spatial_layers(num_layers, in_dim, hidden_dim, skip_layer_idx, out_dim) -> nn.ModuleList:
    layers: List[nn.Module] = [_linear_block(in_dim, hidden_dim)]
    for i in range(1, num_layers - 1):
        layers.append(_SkipLinear(hidden_dim, in_dim, hidden_dim) if i == skip_layer_idx
                      else _linear_block(hidden_dim, hidden_dim))
    layers.append(_linear_block(hidden_dim, out_dim))
    return nn.ModuleList(layers)


def _spatial_params(layers, skip_layer_idx) -> List[torch.Tensor]:
    out = []
    for i, layer in enumerate(layers):
        if i == skip_layer_idx:
            out += [layer.Lx.weight, layer.Lx.bias, layer.Ly.weight, layer.Ly.bias]
        else:
            out += [layer[0].weight, layer[0].bias]
    return out


class TextureField(_FusedFieldBase):
    """6-layer/128-wide ReLU MLP with an input skip connection, RGB sigmoid head (model.py:12-112)."""

    def __init__(self, num_layers, in_dim, hidden_dim, skip_layer_idx, input_feature_embed=None, embed_dim=None,
                 embed_include_input=True, embed_std=1., return_rgb=True, out_dim=RGB_COLOR_DIM, batchnorm=False,
                 activation=nn.ReLU, table_dtype: str = "f32"):
        super().__init__()
        assert num_layers > 2 and 0 < skip_layer_idx and skip_layer_idx < num_layers - 1   # model.py:27
        _check_common(hidden_dim, input_feature_embed, batchnorm, activation)
        if return_rgb and out_dim != RGB_COLOR_DIM:
            raise NotImplementedError("an RGB head has 3 outputs (model.py:9)")
        self.skip_layer_idx = skip_layer_idx
        self.return_rgb = return_rgb
        self.cfg = FieldConfig(k=in_dim, num_layers=num_layers, skip=skip_layer_idx, hidden=hidden_dim)
        self.table_dtype = table_dtype
        self.layers = _spatial_layers(num_layers, in_dim, hidden_dim, skip_layer_idx, out_dim)
        self._init_cache()

    # ---- parameters in the flat / state_dict order the C ABI expects -------------------------------
    def _ordered_params(self) -> List[torch.Tensor]:
        return _spatial_params(self.layers, self.skip_layer_idx)

    def forward(self, batch):
        if not self.return_rgb:
            raise NotImplementedError("a spatial MLP with a bottleneck head only runs fused inside "
                                      "TextureFieldWithViewDependency (model.py:156-157)")
        return super().forward(batch)


class _FourierBands(nn.Module):
    """Holds the reference's non-persistent ``freq_bands`` buffer (layers.py:11-19) so module structure matches."""

    def __init__(self, k, include_input=True):
        super().__init__()
        self.register_buffer("freq_bands", 2 ** torch.arange(0, k) * torch.pi, persistent=False)
        self.include_input = include_input


class TextureFieldWithViewDependency(_FusedFieldBase):
    """Spatial MLP -> bottleneck, Fourier-embedded view direction, 2-layer directional MLP (model.py:123-191);
    the whole thing runs inside the fused kernels."""

    def __init__(self, num_layers, in_dim, hidden_dim, skip_layer_idx, bottleneck_vec_dim, in_dim_view_dir,
                 include_view_dir, view_dir_embedding_size, directional_hidden_dim, input_feature_embed=None,
                 embed_dim=None, embed_include_input=True, embed_std=1., face_normals=None,
                 view_dir_strategy="intrinsic", batchnorm=False, activation=nn.ReLU, table_dtype: str = "f32"):
        super().__init__()
        assert num_layers > 2 and 0 < skip_layer_idx and skip_layer_idx < num_layers - 1   # model.py:27
        _check_common(hidden_dim, input_feature_embed, batchnorm, activation)
        if view_dir_strategy not in ("intrinsic", "extrinsic"):
            raise RuntimeError("Unknown viewing direction strategy.")                        # model.py:186
        if in_dim_view_dir != (1 if view_dir_strategy == "intrinsic" else 3):
            raise ValueError(f"in_dim_view_dir={in_dim_view_dir} does not match the {view_dir_strategy} strategy "
                             "(an angle has 1 dim, a direction 3)")
        self.view_dir_strategy = view_dir_strategy
        self.skip_layer_idx = skip_layer_idx
        self.table_dtype = table_dtype
        if face_normals is not None:
            self.register_buffer("face_normals", face_normals.to(torch.float32).contiguous(), persistent=False)  # model.py:146
        self.spatial_mlp = TextureField(num_layers, in_dim, hidden_dim, skip_layer_idx,
                                        input_feature_embed=input_feature_embed, return_rgb=False,
                                        out_dim=bottleneck_vec_dim, batchnorm=batchnorm, activation=activation)
        self.embedding = _FourierBands(view_dir_embedding_size, include_input=include_view_dir)
        embedding_size = in_dim_view_dir * view_dir_embedding_size * 2 + (in_dim_view_dir if include_view_dir else 0)
        self.directional_mlp = nn.Sequential(nn.Linear(bottleneck_vec_dim + embedding_size, directional_hidden_dim),
                                             activation(), nn.Linear(directional_hidden_dim, RGB_COLOR_DIM), nn.Sigmoid())
        self.cfg = FieldConfig(k=in_dim, num_layers=num_layers, skip=skip_layer_idx, hidden=hidden_dim,
                               view=view_dir_strategy, view_embed=view_dir_embedding_size,
                               view_include=bool(include_view_dir), view_bneck=bottleneck_vec_dim,
                               view_hidden=directional_hidden_dim)
        self._init_cache()

    def _ordered_params(self) -> List[torch.Tensor]:
        return _spatial_params(self.spatial_mlp.layers, self.skip_layer_idx) + [
            self.directional_mlp[0].weight, self.directional_mlp[0].bias,
            self.directional_mlp[2].weight, self.directional_mlp[2].bias]

    def _view_inputs(self, batch):
        dirs = batch["unit_ray_dirs"].contiguous()
        if self.view_dir_strategy == "extrinsic":
            return dirs, None, None
        if getattr(self, "face_normals", None) is None:
            raise RuntimeError("the intrinsic view strategy needs face_normals (make_model(..., mesh=mesh))")
        return dirs, batch["hit_face_idxs"].contiguous(), self.face_normals


def init_weights(m):
    if isinstance(m, nn.Linear):
        torch.nn.init.xavier_uniform_(m.weight.data)   # model.py:194-196


def make_model(model_config: dict, mesh=None) -> nn.Module:
    """Same contract as the reference's make_model (model.py:199-258) for the intrinsic path."""
    view_cfg = model_config.get("view_dependence")
    feature_strategy = model_config.get("feature_strategy", "efuncs")
    if model_config.get("type") == "neutex":
        raise NotImplementedError("the NeuTex baseline is a different model and is out of scope of the B200 path")
    if feature_strategy != "efuncs":
        raise NotImplementedError(f"feature_strategy={feature_strategy!r}: only 'efuncs' is implemented")
    k = model_config["k"]
    in_dim = k if isinstance(k, int) else len(k)                       # model.py:211-215
    act = model_config.get("activation", "relu")
    if act != "relu":
        raise NotImplementedError(f"Activation function {act} is not implemented by the B200 kernels.")
    common = dict(input_feature_embed=feature_strategy, batchnorm=model_config.get("batchnorm", False),
                  table_dtype=model_config.get("table_dtype", "f32"))
    if view_cfg is None:
        model = TextureField(model_config["num_layers"], in_dim, model_config["mlp_hidden_dim"],
                             model_config["skip_layer_idx"], **common)
    else:
        assert mesh is not None                                          # model.py:237
        import numpy as np
        face_normals = torch.from_numpy(np.asarray(mesh.face_normals).copy()).to(dtype=torch.float32)
        model = TextureFieldWithViewDependency(
            model_config["num_layers"], in_dim, model_config["mlp_hidden_dim"], model_config["skip_layer_idx"],
            view_cfg["bottleneck_vec_dim"], view_cfg["in_dim_view_dir"], view_cfg["include_view_dir"],
            view_cfg["embed_size"], view_cfg["directional_hidden_dim"], face_normals=face_normals,
            view_dir_strategy=view_cfg["strategy"], **common)
    model.apply(init_weights)                                           # model.py:257
    return model


def load_trained_model(model_config, weights_path, device, mesh=None):
    """utils.py:22-29: build the model of ``config["model"]`` and load a checkpoint written by the reference
    (either the bare ``state_dict`` or the trainer's ``{"model_state_dict": ...}`` wrapper)."""
    model = make_model(model_config, mesh=mesh)
    data = torch.load(weights_path, map_location="cpu")
    model.load_state_dict(data["model_state_dict"] if "model_state_dict" in data else data)
    return model.to(device)
