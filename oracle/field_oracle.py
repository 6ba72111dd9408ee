"""CPU oracle for the intrinsic-neural-field evaluation path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 torch-on-CPU arithmetic, the algorithm of the one hot
path of tum-vision/intrinsic-neural-fields that this repository re-implements in CUDA:

    gather 3 eigenfunction rows per ray -> barycentric interpolation -> MLP (+ optional
    view-direction branch) -> RGB, and the L1-loss training backward pass.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
arm may import it.  The product package (``intrinsic-neural-fields_b200``) never does: it
has no CPU fallback and fails loudly when the CUDA library is missing.

Parity pinning: the reference ships no tests and no golden vectors.
The oracle is therefore pinned against outputs of the *live reference code* imported from
``/root/reference`` in the build container: ``tests/golden/make_golden.py`` (committed)
generated ``tests/golden/*.npz`` and ``tests/test_oracle.py`` checks this file against
them bit-for-bit / to 1e-6.  Every function cites the reference lines it follows
(paths relative to the reference repository root).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

RGB_COLOR_DIM = 3  # model.py:9


# --------------------------------------------------------------------------------------
# a1: gather + barycentric interpolation
# --------------------------------------------------------------------------------------
def gather_interpolate(E: torch.Tensor, vids: torch.Tensor, bary: torch.Tensor) -> torch.Tensor:
    """x[b,:] = sum_j bary[b,j] * E[vids[b,j],:]            (mesh.py:313-324)

    The reference gathers the (3B,k) rows, reshapes to (B,3,k) and contracts with a bmm
    of the (B,1,3) barycentrics; the same three-term fp32 sum is restated here.
    """
    B = vids.shape[0]
    rows = E[vids.reshape(-1).long()].reshape(B, 3, -1)          # mesh.py:319-320
    return torch.bmm(bary.unsqueeze(1), rows).squeeze(1)         # mesh.py:323-324


def gather_interpolate_batched(E, vids, bary, chunk: int = 1 << 18) -> torch.Tensor:
    """Chunked variant used by the renderer (mesh.py:327-339)."""
    out = torch.zeros((vids.shape[0], E.shape[1]), dtype=torch.float32)
    for lo in range(0, vids.shape[0], chunk):
        hi = min(vids.shape[0], lo + chunk)
        out[lo:hi] = gather_interpolate(E, vids[lo:hi], bary[lo:hi])
    return out


def rescale_standard(table: np.ndarray) -> np.ndarray:
    """'standard' rescale of the eigenfunction table (mesh.py:99-102)."""
    return table / (np.max(table, axis=0, keepdims=True) - np.min(table, axis=0, keepdims=True))


# --------------------------------------------------------------------------------------
# a7: Fourier feature encoding (log-space bands)
# --------------------------------------------------------------------------------------
def fourier_freq_bands(emb: int) -> torch.Tensor:
    """freq_bands = 2**arange(emb) * pi                      (layers.py:13-14)"""
    return (2 ** torch.arange(0, emb)) * torch.pi


def fourier_feat_enc(x: torch.Tensor, emb: int, include_input: bool) -> torch.Tensor:
    """cat(cos(x*f), sin(x*f)[, x]) with (input-major, freq-minor) ordering (layers.py:21-25)."""
    embed = (x[..., None] * fourier_freq_bands(emb)).view(*x.size()[:-1], -1)
    if include_input:
        return torch.cat((embed.cos(), embed.sin(), x), dim=-1)
    return torch.cat((embed.cos(), embed.sin()), dim=-1)


# --------------------------------------------------------------------------------------
# a4 / a5: spatial MLP
# --------------------------------------------------------------------------------------
def spatial_param_names(num_layers: int, skip: int, prefix: str = "") -> Sequence[str]:
    """state_dict / parameters() order of TextureField (model.py:42-96, layers.py:50-58)."""
    names = []
    for i in range(num_layers):
        if i == skip:
            names += [f"{prefix}layers.{i}.Lx.weight", f"{prefix}layers.{i}.Lx.bias",
                      f"{prefix}layers.{i}.Ly.weight", f"{prefix}layers.{i}.Ly.bias"]
        else:
            names += [f"{prefix}layers.{i}.0.weight", f"{prefix}layers.{i}.0.bias"]
    return names


def texture_field_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, num_layers: int, skip: int,
                          return_rgb: bool = True, prefix: str = "") -> torch.Tensor:
    """TextureField.forward on the interpolated embedding x (model.py:98-112).

    layer 0: ReLU(W0 x + b0); hidden: ReLU(W h + b); skip layer: ReLU(Lx(h) + Ly(x))
    (layers.py:60-62); last: Sigmoid (return_rgb) or ReLU (model.py:89-94).
    """
    assert num_layers > 2 and 0 < skip < num_layers - 1          # model.py:27
    res = x
    for i in range(num_layers):
        if i == skip:
            res = F.linear(res, p[f"{prefix}layers.{i}.Lx.weight"], p[f"{prefix}layers.{i}.Lx.bias"]) + \
                  F.linear(x, p[f"{prefix}layers.{i}.Ly.weight"], p[f"{prefix}layers.{i}.Ly.bias"])
            res = torch.relu(res)
        else:
            res = F.linear(res, p[f"{prefix}layers.{i}.0.weight"], p[f"{prefix}layers.{i}.0.bias"])
            if i == num_layers - 1:
                res = torch.sigmoid(res) if return_rgb else torch.relu(res)
            else:
                res = torch.relu(res)
    return res


def relu_margin_per_ray(p: Dict[str, torch.Tensor], x: torch.Tensor, num_layers: int, skip: int) -> torch.Tensor:
    """min over every ReLU unit of |z| / rms(z of its layer), per ray, for the plain TextureField (model.py:98-112).

    A reduced-precision evaluation moves every pre-activation by a few 1e-4 of its layer's scale; a unit whose |z| is
    below that can land on the other side of the ReLU kink, which changes its mask and with it the ray's whole
    contribution to the gradients.  Rays whose margin is well above the rounding noise have identical masks in every
    implementation, so gradients over THOSE rays isolate the arithmetic from the kink noise."""
    res = x
    margin = torch.full((x.shape[0],), float("inf"))
    for i in range(num_layers - 1):
        if i == skip:
            z = F.linear(res, p[f"layers.{i}.Lx.weight"], p[f"layers.{i}.Lx.bias"]) + \
                F.linear(x, p[f"layers.{i}.Ly.weight"], p[f"layers.{i}.Ly.bias"])
        else:
            z = F.linear(res, p[f"layers.{i}.0.weight"], p[f"layers.{i}.0.bias"])
        margin = torch.minimum(margin, (z.abs() / z.pow(2).mean().sqrt()).min(dim=1).values)
        res = torch.relu(z)
    return margin


# --------------------------------------------------------------------------------------
# a6: view-direction branch
# --------------------------------------------------------------------------------------
def angle_between(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """acos(cosine_similarity(a,b))                           (model.py:115-120)"""
    return torch.acos(F.cosine_similarity(a, b, dim=-1))


def view_embedding(unit_ray_dirs, hit_face_idxs, face_normals, strategy: str, emb: int,
                   include_view_dir: bool) -> torch.Tensor:
    """_get_embedded_view_dir (model.py:175-186)."""
    if strategy == "intrinsic":
        n = face_normals[hit_face_idxs.long()]
        ang = angle_between(-unit_ray_dirs, n)
        return fourier_feat_enc(ang.unsqueeze(-1), emb, include_view_dir)
    if strategy == "extrinsic":
        return fourier_feat_enc(unit_ray_dirs, emb, include_view_dir)
    raise RuntimeError("Unknown viewing direction strategy.")


def view_dependent_forward(p, x, unit_ray_dirs, hit_face_idxs, face_normals, *, num_layers, skip,
                           strategy, emb, include_view_dir) -> torch.Tensor:
    """TextureFieldWithViewDependency.forward (model.py:188-191, 168-173)."""
    z = texture_field_forward(p, x, num_layers, skip, return_rgb=False, prefix="spatial_mlp.")
    e = view_embedding(unit_ray_dirs, hit_face_idxs, face_normals, strategy, emb, include_view_dir)
    h = torch.relu(F.linear(torch.cat((z, e), dim=-1), p["directional_mlp.0.weight"],
                            p["directional_mlp.0.bias"]))
    return torch.sigmoid(F.linear(h, p["directional_mlp.2.weight"], p["directional_mlp.2.bias"]))


# --------------------------------------------------------------------------------------
# a8: parameter construction (shapes, names, init)
# --------------------------------------------------------------------------------------
def _linear_init(out_f: int, in_f: int, gen: torch.Generator):
    """nn.Linear default bias init + Xavier-uniform weight (model.py:194-196,257)."""
    w = torch.empty(out_f, in_f)
    bound_w = math.sqrt(6.0 / (in_f + out_f))
    w.uniform_(-bound_w, bound_w, generator=gen)
    b = torch.empty(out_f)
    bound_b = 1.0 / math.sqrt(in_f)
    b.uniform_(-bound_b, bound_b, generator=gen)
    return w, b


def make_params(k: int, hidden: int = 128, num_layers: int = 6, skip: int = 3, out_dim: int = 3,
                view: Optional[dict] = None, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Parameter dict with the reference's state_dict keys/shapes (model.py:42-96,134-173).

    The random stream differs from torch's module construction order; parity tests always
    copy one set of parameters into both implementations, so only shapes/ranges matter.
    """
    gen = torch.Generator().manual_seed(seed)
    prefix = "spatial_mlp." if view is not None else ""
    last_out = view["bottleneck_vec_dim"] if view is not None else out_dim
    p: Dict[str, torch.Tensor] = {}
    for i in range(num_layers):
        in_f = k if i == 0 else hidden
        out_f = last_out if i == num_layers - 1 else hidden
        if i == skip:
            p[f"{prefix}layers.{i}.Lx.weight"], p[f"{prefix}layers.{i}.Lx.bias"] = _linear_init(hidden, hidden, gen)
            p[f"{prefix}layers.{i}.Ly.weight"], p[f"{prefix}layers.{i}.Ly.bias"] = _linear_init(hidden, k, gen)
        else:
            p[f"{prefix}layers.{i}.0.weight"], p[f"{prefix}layers.{i}.0.bias"] = _linear_init(out_f, in_f, gen)
    if view is not None:
        d_in = view["in_dim_view_dir"] * view["embed_size"] * 2
        if view["include_view_dir"]:
            d_in += view["in_dim_view_dir"]                       # model.py:163-166
        dh = view["directional_hidden_dim"]
        p["directional_mlp.0.weight"], p["directional_mlp.0.bias"] = _linear_init(dh, last_out + d_in, gen)
        p["directional_mlp.2.weight"], p["directional_mlp.2.bias"] = _linear_init(RGB_COLOR_DIM, dh, gen)
    return p


# --------------------------------------------------------------------------------------
# whole path + a9 train step
# --------------------------------------------------------------------------------------
def field_forward(p, E, vids, bary, *, num_layers=6, skip=3, view: Optional[dict] = None,
                  unit_ray_dirs=None, hit_face_idxs=None, face_normals=None) -> torch.Tensor:
    """Gather-interpolate then the configured model: RayDataLoader.__next__ + model(batch)
    (ray_dataloader.py:129-133, trainer.py:75)."""
    x = gather_interpolate(E, vids, bary)
    if view is None:
        return texture_field_forward(p, x, num_layers, skip, return_rgb=True)
    return view_dependent_forward(p, x, unit_ray_dirs, hit_face_idxs, face_normals,
                                  num_layers=num_layers, skip=skip, strategy=view["strategy"],
                                  emb=view["embed_size"], include_view_dir=view["include_view_dir"])


def bake_pred_rgbs(p, features: np.ndarray, faces_index_efs: np.ndarray, barycentric_coords: np.ndarray,
                   num_layers: int = 6, skip: int = 3, chunk: int = 1 << 15) -> np.ndarray:
    """Texel prediction of UV baking (bake_texture_field.py:296-319): numpy gather of the three rows, an einsum
    with the barycentric weights at numpy's promoted width (float64 when the weights are float64), a cast of the
    embedding to float32, the MLP per 2^15-texel chunk."""
    out = []
    n = faces_index_efs.shape[0]
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        rows = features[faces_index_efs[lo:hi].reshape(-1)].reshape(hi - lo, 3, -1)
        x = np.einsum("bij,bi->bj", rows, barycentric_coords[lo:hi]).astype(np.float32)
        with torch.no_grad():
            out.append(texture_field_forward(p, torch.from_numpy(x), num_layers, skip).numpy())
    return np.concatenate(out, axis=0) if out else np.zeros((0, 3), np.float32)


def l1_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean |pred-target| over B*3 elements (config.py:117-118)."""
    return F.l1_loss(pred, target)


def loss_and_grads(p, E, vids, bary, target, **kw):
    """Forward, L1 loss, autograd backward as in Trainer._train_step (trainer.py:71-84).

    Returns (loss float, pred (B,3), {name: grad}).  No gradient flows into E/bary
    (none requires grad in the reference either).
    """
    q = {n: t.detach().clone().requires_grad_(True) for n, t in p.items()}
    pred = field_forward(q, E, vids, bary, **kw)
    loss = l1_loss(pred, target)
    loss.backward()
    return float(loss.item()), pred.detach(), {n: t.grad.detach() for n, t in q.items()}


def manual_grads_texture_field(p, E, vids, bary, target, num_layers=6, skip=3):
    """Hand-derived backward of the spatial MLP under L1 loss, in float64 numpy.

    Independent restatement of what autograd does for trainer.py:76-81; used to cross-check
    ``loss_and_grads`` and to document the formulas the CUDA kernels implement:
        dy = sign(y-t)/(3B); dz_L = dy*y*(1-y); dW = dz^T h_prev; db = sum_r dz;
        dh_prev = dz W; dz_prev = dh_prev * (h_prev>0); skip layer feeds dz into Lx and Ly.
    """
    P = {n: t.double().numpy() for n, t in p.items()}
    x = gather_interpolate(E, vids, bary).double().numpy()
    t = target.double().numpy()
    hs = [x]
    h = x
    for i in range(num_layers):
        if i == skip:
            z = h @ P[f"layers.{i}.Lx.weight"].T + P[f"layers.{i}.Lx.bias"] + \
                x @ P[f"layers.{i}.Ly.weight"].T + P[f"layers.{i}.Ly.bias"]
        else:
            z = h @ P[f"layers.{i}.0.weight"].T + P[f"layers.{i}.0.bias"]
        h = 1.0 / (1.0 + np.exp(-z)) if i == num_layers - 1 else np.maximum(z, 0.0)
        hs.append(h)
    y = hs[-1]
    B = y.shape[0]
    loss = np.abs(y - t).mean()
    dz = np.sign(y - t) / (3.0 * B) * y * (1.0 - y)
    g = {}
    for i in reversed(range(num_layers)):
        hin = hs[i]
        if i == skip:
            g[f"layers.{i}.Lx.weight"] = dz.T @ hin
            g[f"layers.{i}.Lx.bias"] = dz.sum(0)
            g[f"layers.{i}.Ly.weight"] = dz.T @ x
            g[f"layers.{i}.Ly.bias"] = dz.sum(0)
            dh = dz @ P[f"layers.{i}.Lx.weight"]
        else:
            g[f"layers.{i}.0.weight"] = dz.T @ hin
            g[f"layers.{i}.0.bias"] = dz.sum(0)
            dh = dz @ P[f"layers.{i}.0.weight"]
        if i > 0:
            dz = dh * (hin > 0)
    return loss, y, g


def _r16(a: np.ndarray) -> np.ndarray:
    """Round to fp16 and back (the operand precision of the CUDA path's tensor-core GEMMs)."""
    return a.astype(np.float16).astype(np.float64)


def emulated_fp16_grads(p, E, vids, bary, target, num_layers=6, skip=3, dz_scale=32.0, table_f16=False,
                        view: Optional[dict] = None, unit_ray_dirs=None, hit_face_idxs=None, face_normals=None,
                        vertex_path: bool = False, proj=None):
    """The reference's forward/backward (model.py:98-112,188-191, trainer.py:76-81) restated in float64 with the
    CUDA path's rounding points: table (optionally), interpolated embedding, weights, hidden activations, the
    Fourier view embedding and backward signals are rounded to fp16 before each matrix product; accumulation
    is exact.

    Purpose: ReLU masks of a reduced-precision forward differ from the fp32 reference wherever |z| is
    below the rounding noise, which alone moves each gradient tensor by a few percent in Frobenius
    norm.  Against THIS restatement the masks agree, so the kernels' arithmetic can be checked tightly.
    Returns (loss, pred, grads) with grads of the mean L1 loss.

    ``vertex_path`` selects the rounding points of the per-vertex evaluation of the two K = k contractions
    (csrc/vertex_project.cu, field_rays.cu, vertex_grad.cu): table rounded to fp16 once (the operand image),
    P = fp16(E16 [W0;Wy]16^T) per vertex, [z0|zy] interpolated from P in fp32, zy rounded to fp16 before it is added in
    the skip layer; backward G[v] = fp16(sum bary * dz) and [dW0;dWy] = G^T E16.  ``proj`` (V,256) replaces the
    restated P by the one the device computed (its fp32 accumulation can land on the other side of an fp16 rounding
    boundary in a few elements, which flips ReLU masks downstream in deep configurations).
    """
    P = {n: t.double().numpy() for n, t in p.items()}
    Ed = E.double().numpy()
    if table_f16 or vertex_path:
        Ed = _r16(Ed)
    b = bary.numpy().astype(np.float32)
    rows = Ed[vids.numpy().reshape(-1)].reshape(vids.shape[0], 3, -1)
    # fp32 interpolation (w0*e0, then two fma) as the producer warps do, then one fp16 rounding
    x32 = (b[:, 0:1] * rows[:, 0].astype(np.float32)).astype(np.float32)
    x32 = (b[:, 1:2].astype(np.float64) * rows[:, 1] + x32).astype(np.float32)
    x32 = (b[:, 2:3].astype(np.float64) * rows[:, 2] + x32).astype(np.float32)
    x = _r16(x32)
    t = target.double().numpy()
    L = num_layers
    pre = "spatial_mlp." if view is not None else ""

    # MMA layers 0..nb: (weight name, bias names); the last spatial layer is an MMA layer only with the view branch
    def names(i):
        if i == skip:
            return f"{pre}layers.{i}.Lx.weight", [f"{pre}layers.{i}.Lx.bias", f"{pre}layers.{i}.Ly.bias"]
        return f"{pre}layers.{i}.0.weight", [f"{pre}layers.{i}.0.bias"]

    n_spatial = L if view is not None else L - 1
    def bias(i):
        wn, bn = names(i)
        acc = P[bn[0]].astype(np.float32)
        for extra in bn[1:]:
            acc = (acc + P[extra].astype(np.float32)).astype(np.float32)
        return acc.astype(np.float64)

    vid_np = vids.numpy()
    if vertex_path:
        W0n, _ = names(0)
        Pz0 = _r16(Ed @ _r16(P[W0n]).T)                                   # (V,128) fp16 projected rows
        Pzy = _r16(Ed @ _r16(P[f"{pre}layers.{skip}.Ly.weight"]).T)
        if proj is not None:
            proj = np.asarray(proj, dtype=np.float64)
            Pz0, Pzy = proj[:, :Pz0.shape[1]], proj[:, Pz0.shape[1]:]

        def interp(T):   # fp32: w0*t0, then two fma
            a = (b[:, 0:1] * T[vid_np[:, 0]].astype(np.float32)).astype(np.float32)
            a = (b[:, 1:2].astype(np.float64) * T[vid_np[:, 1]] + a).astype(np.float32)
            return (b[:, 2:3].astype(np.float64) * T[vid_np[:, 2]] + a).astype(np.float32)
        z0_vp = interp(Pz0).astype(np.float64)
        zy_vp = _r16(interp(Pzy))

    hs, masks = [], []
    h16 = x
    for i in range(n_spatial):
        wn, _ = names(i)
        if vertex_path and i == 0:
            z = (z0_vp.astype(np.float32) + bias(0).astype(np.float32)).astype(np.float64)
        else:
            z = h16 @ _r16(P[wn]).T + bias(i)
        if i == skip:
            z = z + (zy_vp if vertex_path else x @ _r16(P[f"{pre}layers.{i}.Ly.weight"]).T)
        masks.append(z > 0)
        h = np.maximum(z, 0.0)
        hs.append(h)
        h16 = _r16(h)
    e16 = None
    if view is not None:
        e = view_embedding(unit_ray_dirs, hit_face_idxs, face_normals, view["strategy"], view["embed_size"],
                           view["include_view_dir"]).double().numpy()
        e16 = _r16(e)
        bn = view["bottleneck_vec_dim"]
        Wd1 = P["directional_mlp.0.weight"]
        z = h16 @ _r16(Wd1[:, :bn]).T + e16 @ _r16(Wd1[:, bn:]).T + P["directional_mlp.0.bias"]
        masks.append(z > 0)
        hs.append(np.maximum(z, 0.0))
        wl_n, bl_n = "directional_mlp.2.weight", "directional_mlp.2.bias"
    else:
        wl_n, bl_n = f"layers.{L-1}.0.weight", f"layers.{L-1}.0.bias"
    z = hs[-1] @ P[wl_n].T + P[bl_n]                 # output layer runs in fp32 on CUDA cores
    y = 1.0 / (1.0 + np.exp(-z))
    B = y.shape[0]
    loss = np.abs(y - t).mean()
    dzl = np.sign(y - t) * y * (1.0 - y) * dz_scale
    g = {}
    g[wl_n] = _r16(dzl).T @ _r16(hs[-1])             # tensor-core product of the stashed fp16 tiles
    g[bl_n] = dzl.sum(0)
    nb = len(hs) - 1
    dz = _r16((dzl @ P[wl_n]) * masks[nb])
    if view is not None:
        bn = view["bottleneck_vec_dim"]
        Wd1 = P["directional_mlp.0.weight"]
        g["directional_mlp.0.weight"] = np.concatenate([dz.T @ _r16(hs[nb - 1]), dz.T @ e16], axis=1)
        g["directional_mlp.0.bias"] = dz.sum(0)
        dz = _r16((dz @ _r16(Wd1[:, :bn])) * masks[nb - 1])
    def vertex_grad(dzr):
        # G[v] = fp16(sum over incident (ray, corner) of bary * dz[ray]); dW = G^T E16
        G = np.zeros((Ed.shape[0], dzr.shape[1]))
        for j in range(3):
            np.add.at(G, vid_np[:, j], b[:, j:j + 1].astype(np.float64) * dzr)
        return _r16(G).T @ Ed

    for i in range(n_spatial - 1, -1, -1):
        hin = x if i == 0 else _r16(hs[i - 1])
        wn, bns = names(i)
        g[wn] = vertex_grad(dz) if (vertex_path and i == 0) else dz.T @ hin
        for bname in bns:
            g[bname] = dz.sum(0)
        if i == skip:
            g[f"{pre}layers.{i}.Ly.weight"] = vertex_grad(dz) if vertex_path else dz.T @ x
        if i > 0:
            dz = _r16((dz @ _r16(P[wn])) * masks[i - 1])
    s = 1.0 / (3.0 * B * dz_scale)
    return loss, y, {n: v * s for n, v in g.items()}


def adam_train_steps(p, E, vids, bary, target, batch_index_lists, lr=1e-4, **kw):
    """Fixed-step Adam training as trainer.py:71-84 + config.py:108 (Adam(lr), defaults).

    ``batch_index_lists`` is an explicit list of index tensors so CPU and CUDA runs see the
    same batches (CPU/CUDA randperm streams differ).
    Returns (final params, list of per-step losses).
    """
    q = {n: t.detach().clone().requires_grad_(True) for n, t in p.items()}
    opt = torch.optim.Adam(list(q.values()), lr=lr)
    losses = []
    for idx in batch_index_lists:
        pred = field_forward(q, E, vids[idx], bary[idx],
                             **{k_: (v[idx] if k_ in ("unit_ray_dirs", "hit_face_idxs") and v is not None else v)
                                for k_, v in kw.items()})
        loss = l1_loss(pred, target[idx])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.item()))
    return {n: t.detach() for n, t in q.items()}, losses


def epoch_psnr(sum_sq_err: float, num_rays: int) -> float:
    """-10 log10(sum of squared error / number of rays)  (trainer.py:254,263; evaluation_metrics.py:25-26)."""
    return float(-10.0 * np.log10(sum_sq_err / num_rays))
