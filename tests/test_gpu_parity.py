"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
produced by the live reference.  Tolerances are BASELINE.json's: per-ray RGB <= 2e-2 max abs and
<= 1e-3 mean abs; gradients are compared by relative Frobenius error."""
import numpy as np
import pytest
import torch

from conftest import load_golden, oracle_kwargs
from oracle import field_oracle as O

import inf_b200
from inf_b200 import ops

pytestmark = pytest.mark.gpu

RGB_MAX, RGB_MEAN = 2e-2, 1e-3      # BASELINE.json north_star tolerances
# Gradients.  A reduced-precision forward flips the ReLU mask wherever |z| is below the rounding noise;
# those few elements alone move every gradient tensor by 1-4 % in Frobenius norm against the fp32
# reference (measured; identical for tf32).  So: a loose bound + cosine against the reference, and a
# tight bound against the oracle's restatement with the kernels' rounding points (masks then agree).
GRAD_REL = 6e-2
GRAD_COS = 0.997
GRAD_REL_EMULATED = 1e-2
GRAD_REL_VIEW, GRAD_COS_VIEW = 0.2, 0.98

DEV = "cuda:0"


@pytest.fixture(autouse=True, params=["ray", "vertex"])
def eval_path(request):
    """Every case of this module runs on both evaluations of the two K = k contractions (InfDesc.path): per ray
    (field_main.cu) and per vertex (vertex_project.cu / field_rays.cu / vertex_grad.cu)."""
    old = ops.DEFAULT_PATH
    ops.DEFAULT_PATH = request.param
    yield request.param
    ops.DEFAULT_PATH = old


def _model_from(case, table_dtype="f32"):
    cfg = dict(k=case["k"], num_layers=case["num_layers"], mlp_hidden_dim=128, skip_layer_idx=case["skip"])
    mesh = None
    if case.get("view") is not None:
        import types
        cfg["view_dependence"] = case["view"]
        mesh = types.SimpleNamespace(face_normals=case["face_normals"].numpy())
    m = inf_b200.make_model(cfg, mesh=mesh)
    m.load_state_dict(case["params"])
    return m.to(DEV)


def _batch(case, table_dtype="f32", idx_dtype=torch.int64):
    tab = inf_b200.EigenTable.from_dense(case["table"], device=DEV, dtype=table_dtype)
    b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": case["vids"].to(DEV).to(idx_dtype),
         "barycentric_coords": case["bary"].to(DEV), "expected_rgbs": case["target"].to(DEV)}
    if case.get("view") is not None:
        b["unit_ray_dirs"] = case["unit_ray_dirs"].to(DEV)
        b["hit_face_idxs"] = case["face_idxs"].to(DEV)
    return b


def _check_no_watchdog():
    torch.cuda.synchronize()
    assert inf_b200.kernel_error_word(DEV) == 0


# ---------------------------------------------------------------------------------------------
# tensor-core operand layouts (localises descriptor bugs before the full kernels are judged)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,n,k", [(0, 128, 128), (0, 256, 64), (0, 16, 64), (1, 16, 128), (1, 128, 64),
                                      (2, 128, 128), (2, 256, 64), (2, 64, 128), (3, 128, 128), (3, 64, 64)])
def test_umma_operand_layouts(mode, n, k):
    g = torch.Generator().manual_seed(mode * 100 + n + k)
    A = (torch.randn(128, k, generator=g)).half().float()
    B = (torch.randn(n, k, generator=g)).half().float()
    ref = A.double() @ B.double().T
    a_in = A if mode in (0, 3) else A.T.contiguous()      # MN-major sources are (k,128) / (k,n)
    b_in = B if mode < 2 else B.T.contiguous()
    d = inf_b200.debug_umma(mode, a_in.to(DEV), b_in.to(DEV)).cpu().double()
    assert torch.allclose(d, ref, atol=1e-2, rtol=1e-3), float((d - ref).abs().max())


# ---------------------------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tf_k64", "tf_k1023", "tf_k100_l5s1", "view_intrinsic", "view_extrinsic"])
@pytest.mark.parametrize("table_dtype", ["f32", "f16"])
def test_forward_matches_reference_golden(name, table_dtype):
    c = load_golden(name)
    m = _model_from(c)
    with torch.no_grad():
        pred = m(_batch(c, table_dtype)).cpu()
    _check_no_watchdog()
    err = (pred - c["pred"]).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN, (err.max().item(), err.mean().item())


def test_forward_int32_indices_and_dense_path():
    c = load_golden("tf_k64")
    m = _model_from(c)
    with torch.no_grad():
        p64 = m(_batch(c)).cpu()
        p32 = m(_batch(c, idx_dtype=torch.int32)).cpu()
        pd = m({"eigenfunctions": c["x"].to(DEV)}).cpu()
    _check_no_watchdog()
    assert torch.equal(p64, p32)                     # index dtype is a pure pass-through
    assert (pd - c["pred"]).abs().max().item() <= RGB_MAX
    assert (pd - p64).abs().max().item() <= 2e-3


def test_degenerate_barycentrics_select_vertex_rows():
    # the reference's only self-check uses bary = [1,0,0]: the embedding is E[vids[:,0]]
    c = load_golden("tf_k64")
    m = _model_from(c)
    bary = torch.zeros_like(c["bary"]); bary[:, 0] = 1.0
    b = _batch(c); b["barycentric_coords"] = bary.to(DEV)
    with torch.no_grad():
        pred = m(b).cpu()
    ref = O.texture_field_forward(c["params"], c["table"][c["vids"][:, 0]], c["num_layers"], c["skip"])
    assert (pred - ref).abs().max().item() <= RGB_MAX


@pytest.mark.parametrize("B", [1, 127, 128, 129, 1000])
def test_ragged_batch_sizes(B):
    c = load_golden("tf_k100_l5s1")
    g = torch.Generator().manual_seed(B)
    idx = torch.randint(0, c["vids"].shape[0], (B,), generator=g)
    m = _model_from(c)
    b = _batch(c)
    for key in ("vertex_idxs_of_hit_faces", "barycentric_coords", "expected_rgbs"):
        b[key] = b[key][idx.to(DEV)].contiguous()
    with torch.no_grad():
        pred = m(b).cpu()
    _check_no_watchdog()
    ref = O.field_forward(c["params"], c["table"], c["vids"][idx], c["bary"][idx], num_layers=c["num_layers"], skip=c["skip"])
    err = (pred - ref).abs()
    assert pred.shape == (B, 3)
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN


def test_empty_batch():
    c = load_golden("tf_k64")
    m = _model_from(c)
    b = _batch(c)
    for key in ("vertex_idxs_of_hit_faces", "barycentric_coords", "expected_rgbs"):
        b[key] = b[key][:0].contiguous()
    with torch.no_grad():
        assert m(b).shape == (0, 3)


# ---------------------------------------------------------------------------------------------
# training: fused L1 step and autograd path
# ---------------------------------------------------------------------------------------------
def _check_grads(m, ref_grads, tol=GRAD_REL, cos_min=GRAD_COS):
    worst = 0.0
    for n, p in m.named_parameters():
        ref = ref_grads[n]
        if not torch.is_tensor(ref):
            ref = torch.from_numpy(np.asarray(ref)).float()
        ref = ref.reshape(p.shape).float()
        g = p.grad.detach().cpu()
        rel = (g - ref).norm().item() / (ref.norm().item() + 1e-12)
        cos = (g * ref).sum().item() / (g.norm().item() * ref.norm().item() + 1e-30)
        worst = max(worst, rel)
        assert rel <= tol and cos >= cos_min, (n, rel, cos)
    return worst


def _device_proj(m, tab):
    """On the vertex path the emulation continues from the P the device computed (inf_project_vertices): its fp32
    accumulation lands on the other side of an fp16 rounding boundary in a handful of elements, which is enough to flip
    ReLU masks downstream in deep configurations.  P itself is checked against the restatement to one fp16 ulp."""
    if ops.DEFAULT_PATH != "vertex":
        return None
    return ops.project_vertices(m.cfg, tab, m._packed_weights()).float().cpu().numpy()


def _check_grads_emulated(m, case, table_f16=False, tab=None):
    kw = {}
    if case.get("view") is not None:
        kw = dict(view=case["view"], unit_ray_dirs=case["unit_ray_dirs"], hit_face_idxs=case["face_idxs"],
                  face_normals=case["face_normals"])
    _, _, g = O.emulated_fp16_grads(case["params"], case["table"], case["vids"], case["bary"], case["target"],
                                    case["num_layers"], case["skip"], table_f16=table_f16,
                                    vertex_path=(ops.DEFAULT_PATH == "vertex"), proj=None if tab is None else _device_proj(m, tab),
                                    **kw)
    return _check_grads(m, g, tol=GRAD_REL_EMULATED, cos_min=0.9999)


@pytest.mark.parametrize("name", ["tf_k64", "tf_k1023", "tf_k100_l5s1", "view_intrinsic", "view_extrinsic"])
@pytest.mark.parametrize("table_dtype", ["f32", "f16"])
def test_fused_l1_step_matches_reference_golden(name, table_dtype):
    c = load_golden(name)
    m = _model_from(c)
    b = _batch(c, table_dtype)
    sums, pred, flat = m.l1_step(b)
    _check_no_watchdog()
    B = c["target"].shape[0]
    loss = sums[0].item() / (3 * B)
    assert abs(loss - c["loss"]) <= 1e-3
    sse_ref = ((c["pred"] - c["target"]) ** 2).sum().item()
    assert abs(sums[1].item() - sse_ref) <= 2e-2 * max(1.0, sse_ref)
    assert (pred.cpu() - c["pred"]).abs().max().item() <= RGB_MAX
    # the deeper view-branch chain on 96-140 rays amplifies the ReLU-mask flips of a reduced-precision forward
    _check_grads(m, c["grads"], tol=GRAD_REL if c["view"] is None else GRAD_REL_VIEW,
                 cos_min=GRAD_COS if c["view"] is None else GRAD_COS_VIEW)
    _check_grads_emulated(m, c, table_f16=(table_dtype == "f16"), tab=b["eigenfunction_table"])


def test_autograd_path_matches_fused_step_and_reference():
    c = load_golden("tf_k64")
    m = _model_from(c)
    b = _batch(c)
    pred = m(b)
    loss = torch.nn.functional.l1_loss(pred, b["expected_rgbs"])     # trainer.py:76
    loss.backward()
    _check_no_watchdog()
    assert abs(loss.item() - c["loss"]) <= 1e-3
    _check_grads(m, c["grads"])
    # L2 loss goes through the same backward kernels with a different upstream gradient
    m2 = _model_from(c)
    torch.nn.functional.mse_loss(m2(b), b["expected_rgbs"]).backward()
    q = {n: t.detach().clone().requires_grad_(True) for n, t in c["params"].items()}
    ref = torch.nn.functional.mse_loss(O.field_forward(q, c["table"], c["vids"], c["bary"], num_layers=6, skip=3), c["target"])
    ref.backward()
    _check_grads(m2, {n: t.grad for n, t in q.items()})


def test_view_branch_padded_widths_and_int32_faces():
    # bottleneck / directional widths below 128 are zero-padded inside; face ids follow the index dtype of the call
    torch.manual_seed(3)
    view = dict(strategy="intrinsic", in_dim_view_dir=1, include_view_dir=False, embed_size=6,
                bottleneck_vec_dim=96, directional_hidden_dim=48)
    V, F_, B, k = 200, 300, 333, 80
    E = torch.randn(V, k) / 4
    normals = torch.nn.functional.normalize(torch.randn(F_, 3), dim=-1)
    fid = torch.randint(0, F_, (B,))
    dirs = torch.nn.functional.normalize(torch.randn(B, 3), dim=-1)
    dirs = torch.where(((-dirs * normals[fid]).sum(-1) < 0)[:, None], -dirs, dirs)
    vids = torch.randint(0, V, (B, 3))
    u = torch.rand(B, 3); bary = u / u.sum(-1, keepdim=True)
    target = torch.rand(B, 3)
    params = O.make_params(k, num_layers=5, skip=2, view=view, seed=11)
    import types
    m = inf_b200.make_model(dict(k=k, num_layers=5, mlp_hidden_dim=128, skip_layer_idx=2, view_dependence=view),
                            mesh=types.SimpleNamespace(face_normals=normals.numpy()))
    m.load_state_dict(params)
    m = m.to(DEV)
    b = {"eigenfunction_table": inf_b200.EigenTable.from_dense(E, device=DEV), "vertex_idxs_of_hit_faces": vids.to(DEV).int(),
         "barycentric_coords": bary.to(DEV), "expected_rgbs": target.to(DEV), "unit_ray_dirs": dirs.to(DEV),
         "hit_face_idxs": fid.to(DEV)}
    sums, pred, flat = m.l1_step(b)
    _check_no_watchdog()
    kw = dict(num_layers=5, skip=2, view=view, unit_ray_dirs=dirs, hit_face_idxs=fid, face_normals=normals)
    loss, pref, grads = O.loss_and_grads(params, E, vids, bary, target, **kw)
    err = (pred.cpu() - pref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN
    assert abs(sums[0].item() / (3 * B) - loss) <= 1e-3
    _check_grads(m, grads)


@pytest.mark.parametrize("k,L,skip", [(1023, 6, 3), (200, 8, 6)])
def test_gradients_match_the_reference_on_rays_away_from_relu_kinks(k, L, skip):
    """The hypothesis behind GRAD_REL (6e-2) tested directly: restricted to rays whose every pre-activation is further
    from zero than the rounding noise -- so that the fp32 reference and the fp16-operand kernels take the SAME ReLU
    masks -- the gradients agree with the REFERENCE fp32 gradients (autograd of the oracle) to 1e-2 per tensor, on both
    evaluation paths.  (On all rays the same comparison gives 1-4 %: tests above.)"""
    g = torch.Generator().manual_seed(k + L)
    V, B = 3000, 16384
    E = torch.randn(V, k, generator=g) / 6
    faces = torch.randint(0, V, (6000, 3), generator=g)
    vids = faces[torch.randint(0, 6000, (B,), generator=g)]
    u = torch.rand(B, 3, generator=g)
    bary = u / u.sum(-1, keepdim=True)
    target = torch.rand(B, 3, generator=g)
    params = O.make_params(k, num_layers=L, skip=skip, seed=L)
    margin = O.relu_margin_per_ray(params, O.gather_interpolate(E, vids, bary), L, skip)
    keep = torch.nonzero(margin > 4e-3).reshape(-1)          # ~10x the fp16 rounding noise of a pre-activation
    assert keep.numel() >= 300, keep.numel()
    vs, bs, ts = vids[keep], bary[keep], target[keep]
    loss, pref, grads = O.loss_and_grads(params, E, vs, bs, ts, num_layers=L, skip=skip)
    m = inf_b200.make_model(dict(k=k, num_layers=L, mlp_hidden_dim=128, skip_layer_idx=skip))
    m.load_state_dict(params)
    m = m.to(DEV)
    b = {"eigenfunction_table": inf_b200.EigenTable.from_dense(E, device=DEV), "vertex_idxs_of_hit_faces": vs.to(DEV),
         "barycentric_coords": bs.to(DEV), "expected_rgbs": ts.to(DEV)}
    sums, pred, flat = m.l1_step(b)
    _check_no_watchdog()
    assert (pred.cpu() - pref).abs().max().item() <= 2e-3
    _check_grads(m, grads, tol=1e-2, cos_min=0.9999)


def test_gradients_are_deterministic():
    c = load_golden("tf_k1023")
    m = _model_from(c)
    b = _batch(c)
    f1 = m.l1_step(b)[2].clone()
    f2 = m.l1_step(b)[2].clone()
    assert torch.equal(f1, f2)


def test_large_batch_linearity_of_gradients():
    # size-independent property at a full-size batch: grads over a batch == sum of grads over its halves
    torch.manual_seed(0)
    V, k, B = 5000, 1023, 8192
    E = torch.randn(V, k) / 8
    m = inf_b200.make_model(dict(k=k, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3)).to(DEV)
    tab = inf_b200.EigenTable.from_dense(E, device=DEV)
    vids = torch.randint(0, V, (B, 3), device=DEV)
    u = torch.rand(B, 3, device=DEV); bary = u / u.sum(-1, keepdim=True)
    tgt = torch.rand(B, 3, device=DEV)
    def run(sl):
        b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": vids[sl].contiguous(),
             "barycentric_coords": bary[sl].contiguous(), "expected_rgbs": tgt[sl].contiguous()}
        sums, pred, flat = m.l1_step(b, global_rays=B)
        return sums.clone(), pred.clone(), flat.clone()
    s_all, p_all, f_all = run(slice(0, B))
    s_a, p_a, f_a = run(slice(0, B // 2))
    s_b, p_b, f_b = run(slice(B // 2, B))
    _check_no_watchdog()
    assert torch.equal(p_all, torch.cat([p_a, p_b]))                     # per-ray results are batch independent
    assert torch.allclose(s_all, s_a + s_b, rtol=1e-4)
    rel = (f_all - (f_a + f_b)).norm().item() / f_all.norm().item()
    assert rel < 1e-4, rel
    # and against the oracle on a subsample of rays (forward) at this size
    idx = torch.arange(0, B, 64)
    ref = O.field_forward({n: p.detach().cpu() for n, p in m.state_dict().items()}, E, vids.cpu()[idx], bary.cpu()[idx])
    err = (p_all.cpu()[idx] - ref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN


# ---------------------------------------------------------------------------------------------
# fixed-step training: PSNR parity with the reference fp32 path (BASELINE.json: within 0.1 dB)
# ---------------------------------------------------------------------------------------------
PSNR_TOL_DB = 0.1


def _training_problem(view, seed=0, V=1500, F_=2500, k=64, N=8192):
    from psnr_protocols import problem
    return problem(dict(view=view, V=V, F=F_, k=k, N=N), seed)


def _psnr_ensemble(name):
    """PSNRs (and final losses) of the fp32 reference under perturbations of its initial weights the size of fp32
    summation-order noise (tests/golden/make_psnr_ensemble.py): the reference's own noise floor for this protocol."""
    import json
    import os
    from conftest import GOLDEN_DIR
    return json.load(open(os.path.join(GOLDEN_DIR, "psnr_ensemble.json")))[name]


def _run_protocol(name, fused_adam=False):
    """Train the protocol's model on the GPU with the committed batch index lists; returns (losses, held-out PSNR)."""
    import types
    from conftest import VIEW_CFGS
    from inf_b200 import trainer
    from psnr_protocols import PROTOCOLS, problem, index_lists
    pr = PROTOCOLS[name]
    view = VIEW_CFGS[pr["view"]] if pr["view"] else None
    E, vids, bary, target, dirs, fid, normals = problem(pr)
    params = O.make_params(pr["k"], num_layers=pr["L"], skip=pr["skip"], view=view, seed=5)
    cfg = dict(k=pr["k"], num_layers=pr["L"], mlp_hidden_dim=128, skip_layer_idx=pr["skip"])
    mesh = None
    if view is not None:
        cfg["view_dependence"] = view
        mesh = types.SimpleNamespace(face_normals=normals.numpy())
    m = inf_b200.make_model(cfg, mesh=mesh)
    m.load_state_dict(params)
    m = m.to(DEV)
    opt = trainer.FusedFieldAdam(m, lr=pr["lr"]) if fused_adam else torch.optim.Adam(m.parameters(), lr=pr["lr"])
    tab = inf_b200.EigenTable.from_dense(E, device=DEV)
    dv, db, dt, dd, df = (t.to(DEV) for t in (vids, bary, target, dirs, fid))

    def batch(sel):
        b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": dv[sel].contiguous(),
             "barycentric_coords": db[sel].contiguous(), "expected_rgbs": dt[sel].contiguous()}
        if view is not None:
            b["unit_ray_dirs"] = dd[sel].contiguous()
            b["hit_face_idxs"] = df[sel].contiguous()
        return b

    losses = []
    for idx in index_lists(pr):
        if fused_adam:
            loss, _ = trainer.fused_train_step_async(m, opt, batch(idx.to(DEV)))
        else:
            loss, _ = trainer.train_step(m, opt, batch(idx.to(DEV)))                 # trainer.py:71-84
        losses.append(loss)
    if fused_adam:
        losses = [float(x) for x in torch.stack(losses).cpu()]
    _check_no_watchdog()
    hold = slice(pr["ntrain"], None)
    with torch.no_grad():
        pg = torch.cat([m(batch(slice(lo, min(lo + 16384, pr["N"])))).cpu() for lo in range(pr["ntrain"], pr["N"], 16384)])
    psnr_gpu = O.epoch_psnr(float(((pg - target[hold]) ** 2).sum()), pg.shape[0])   # trainer.py:254,263
    # per-ray RGB parity AT THE TRAINED WEIGHTS: the fp32 oracle evaluated with the weights the GPU run ended with
    sub = torch.arange(pr["ntrain"], pr["N"], max(1, (pr["N"] - pr["ntrain"]) // 2048))
    kw = dict(num_layers=pr["L"], skip=pr["skip"])
    if view is not None:
        kw.update(view=view, unit_ray_dirs=dirs[sub], hit_face_idxs=fid[sub], face_normals=normals)
    trained = {n: t.detach().cpu() for n, t in m.state_dict().items()}
    with torch.no_grad():
        ref = O.field_forward(trained, E, vids[sub], bary[sub], **kw)
    err = (pg[sub - pr["ntrain"]] - ref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN, (name, err.max().item(), err.mean().item())
    return losses, psnr_gpu


def _assert_psnr_parity(name, losses, psnr_gpu):
    ens = _psnr_ensemble(name)
    lo, hi = min(e["psnr"] for e in ens), max(e["psnr"] for e in ens)
    base = next(e for e in ens if e["eps"] == 0.0)
    assert losses[-1] < 0.8 * losses[0]                                          # it trained
    assert abs(losses[-1] - base["final_loss"]) < 5e-3
    # within 0.1 dB (BASELINE.json) of what the fp32 reference itself produces under fp32-roundoff-sized perturbations
    assert lo - PSNR_TOL_DB <= psnr_gpu <= hi + PSNR_TOL_DB, (name, psnr_gpu, lo, hi, base["psnr"])
    return psnr_gpu - base["psnr"]


@pytest.mark.parametrize("name", ["small_none", "small_view"])
def test_fixed_step_training_psnr_parity(name):
    # lr = the shipped configs' 1e-4 (intrinsic_cat.yaml:33); 400 steps of 512 rays, k = 64.  The plain model's reference
    # ensemble spans 0.11 dB, the view-dependent model's 0.50 dB (tests/golden/psnr_ensemble.json).
    losses, psnr_gpu = _run_protocol(name)
    _assert_psnr_parity(name, losses, psnr_gpu)


@pytest.mark.parametrize("name,fused_adam", [("cat_2000", False), ("cat_2000", True), ("view_2000", True)])
def test_2000_step_training_at_config_scale_psnr_and_rgb_parity(name, fused_adam):
    # configs[1] at its shipped sizes: k = 1023, V = 50 176, batch 4096, lr 1e-4 (intrinsic_cat.yaml), 2000 optimisation
    # steps over identical batch index lists; the view-dependent model alike.  After ~1000 steps
    # the weights have left their initialisation scale -- where SURVEY.md section 7.3 found all-bf16 operands to fail.
    # Both the torch.optim.Adam step and the optimiser fused into the reduction kernel are covered.
    # Measured sensitivity of the view-dependent protocol (scripts/psnr_values.py, psnr_sensitivity.sh, grad_split_check.py): the
    # reference ensemble under 1e-6 / 1e-5 perturbations spans 28.96 .. 29.03 dB, under 1e-4 / 1e-3 perturbations 28.75 .. 29.05 dB
    # ("view_2000_wide" in psnr_ensemble.json, not asserted against); the GPU runs land at 28.95 .. 29.07 dB (both paths, both
    # optimisers), and a different split of the wgrad CTAs -- one-step gradients equal to 9e-8, i.e. fp32 re-association only --
    # moves them to 28.59 .. 28.86 dB.  The 0.1 dB bound therefore holds for the shipped (deterministic) summation order.
    losses, psnr_gpu = _run_protocol(name, fused_adam=fused_adam)
    _assert_psnr_parity(name, losses, psnr_gpu)


# ---------------------------------------------------------------------------------------------
# data path: bit-exact pass-through of the gather inputs, host loaders
# ---------------------------------------------------------------------------------------------
def test_ray_dataloader_passes_gather_inputs_through_bit_exactly():
    from inf_b200.ray_dataloader import RayDataLoader, HostRayLoader
    c = load_golden("tf_k64")
    N = c["vids"].shape[0]
    dl = RayDataLoader(c["table"], "efuncs", c["vids"], c["bary"], c["target"], None, None, batch_size=100, shuffle=False,
                       drop_last=False, device=DEV)
    seen = 0
    for b in dl:
        n = b["expected_rgbs"].shape[0]
        assert torch.equal(b["vertex_idxs_of_hit_faces"].cpu(), c["vids"][seen:seen + n])       # int64 as dataset.py:16
        assert torch.equal(b["barycentric_coords"].cpu(), c["bary"][seen:seen + n])             # f32 bits untouched
        assert torch.equal(b["expected_rgbs"].cpu(), c["target"][seen:seen + n])
        seen += n
    assert seen == N and len(dl) == 3
    # host-resident split, int32 ids as stored on disk (mesh.py:513-516), copied one batch ahead
    tab = inf_b200.EigenTable.from_dense(c["table"], device=DEV)
    hl = HostRayLoader(tab, c["vids"].to(torch.int32), c["bary"], c["target"], batch_size=100, device=DEV)
    m = _model_from(c)
    seen, preds = 0, []
    for b in hl:
        n = b["expected_rgbs"].shape[0]
        assert b["vertex_idxs_of_hit_faces"].dtype == torch.int32
        assert torch.equal(b["vertex_idxs_of_hit_faces"].cpu().long(), c["vids"][seen:seen + n])
        assert torch.equal(b["barycentric_coords"].cpu(), c["bary"][seen:seen + n])
        with torch.no_grad():
            preds.append(m(b).cpu())
        seen += n
    assert seen == N
    assert (torch.cat(preds) - c["pred"]).abs().max().item() <= RGB_MAX


def test_renderer_host_pipeline_matches_device_path_and_canvas():
    from inf_b200.renderer import FieldRenderer
    c = load_golden("tf_k64")
    m = _model_from(c)
    M = c["vids"].shape[0]
    H, W = 20, 20
    r = FieldRenderer(m, c["table"], H=H, W=W, device=DEV)
    a = r.predict_hits(c["vids"], c["bary"]).cpu()
    b = r.predict_hits_host(c["vids"].to(torch.int32), c["bary"], chunk=64)
    assert torch.equal(a, b)                                   # chunking / index width do not change a single bit
    assert (a - c["pred"]).abs().max().item() <= RGB_MAX
    # canvas assembly of renderer.py:119-130: hits scattered by ray index, background elsewhere
    hit = torch.randperm(H * W)[:M]
    img = r.render_hits(c["vids"], c["bary"], hit)
    flat = torch.from_numpy(img).reshape(-1, 3)
    assert torch.equal(flat[hit], a)
    rest = torch.ones(H * W, dtype=torch.bool); rest[hit] = False
    assert torch.all(flat[rest] == 1.0)


def test_graphed_train_step_equals_eager_steps():
    # the CUDA-graph replay of the whole step (pack, forward/backward, wgrad, reduce, Adam) must reproduce eager steps
    from inf_b200 import trainer
    c = load_golden("tf_k64")
    b = _batch(c)
    runs = []
    for graphed in (False, True):
        m = _model_from(c)
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
        gs = trainer.GraphedTrainStep(m, opt)
        losses = []
        for _ in range(5):
            loss, pred = gs.step(b) if graphed else trainer.train_step(m, opt, b)
            losses.append(loss)
        runs.append((losses, [p.detach().clone() for p in m.parameters()], pred.clone()))
    _check_no_watchdog()
    (l0, p0, y0), (l1, p1, y1) = runs
    assert l0 == l1
    assert torch.equal(y0, y1)
    for a, bb in zip(p0, p1):
        assert torch.equal(a, bb)


def test_fused_adam_without_version_bump_is_not_served_stale_weights():
    # torch.optim.Adam(fused=True) updates parameters in place WITHOUT bumping ._version; the operand images must follow
    from inf_b200 import trainer
    c = load_golden("tf_k64")
    b = _batch(c)
    out = []
    for fused in (False, True):
        m = _model_from(c)
        opt = torch.optim.Adam(m.parameters(), lr=1e-2, fused=fused)
        losses = [trainer.train_step(m, opt, b)[0] for _ in range(4)]
        m.eval()
        with torch.no_grad():
            out.append((losses, m(b).cpu()))
    (l0, y0), (l1, y1) = out
    assert l0[0] == l1[0]
    assert max(abs(a - bb) for a, bb in zip(l0, l1)) < 1e-4       # same trajectory (fused vs foreach arithmetic differs by ulps)
    assert l0[-1] < l0[0] and l1[-1] < l1[0] and len(set(l1)) == 4   # stale operand images would freeze the loss
    assert (y0 - y1).abs().max().item() < 1e-2                        # eval after train() -> eval() sees the last update


@pytest.mark.parametrize("name", ["tf_k64", "view_intrinsic"])
def test_fused_adam_matches_torch_adam(name):
    # Adam + operand refresh inside the reduction kernel == torch.optim.Adam + re-pack (config.py:108), step by step
    from inf_b200 import trainer
    c = load_golden(name)
    b = _batch(c)
    m0, m1 = _model_from(c), _model_from(c)
    opt = torch.optim.Adam(m0.parameters(), lr=1e-4)
    fopt = trainer.FusedFieldAdam(m1, lr=1e-4)
    assert [n for n, _ in m1.named_parameters()] == c["param_names"]          # still the reference's state_dict
    for step in range(6):
        l0, p0 = trainer.train_step(m0, opt, b)
        l1, p1 = trainer.fused_train_step(m1, fopt, b)
        assert abs(l0 - l1) < 2e-5, (step, l0, l1)
    _check_no_watchdog()
    assert fopt.step_count == 6
    for (n, a), (_, bb) in zip(m0.named_parameters(), m1.named_parameters()):
        # a few ReLU-kink elements can flip with the 1-ulp differences of the update arithmetic; compare in bulk
        rel = (a - bb).norm().item() / a.norm().item()
        assert rel < 2e-4, (n, rel)
    # optimiser checkpoints are interchangeable with torch.optim.Adam's
    sd, ref_sd = fopt.state_dict(), opt.state_dict()
    assert sd["state"].keys() == ref_sd["state"].keys()
    for i in sd["state"]:
        assert sd["state"][i]["exp_avg"].shape == ref_sd["state"][i]["exp_avg"].shape
        assert float(sd["state"][i]["step"]) == float(ref_sd["state"][i]["step"])
    m1.eval()
    with torch.no_grad():
        ya, yb = m0.eval()(b), m1(b)
    assert (ya - yb).abs().max().item() < 2e-3
    # the flat master buffer IS the parameters
    assert all(p.data_ptr() >= fopt.flat.data_ptr() and p.data_ptr() < fopt.flat.data_ptr() + fopt.flat.numel() * 4
               for p in m1.parameters())


def test_graphed_fused_step_equals_eager_fused_steps():
    from inf_b200 import trainer
    c = load_golden("tf_k64")
    b = _batch(c)
    runs = []
    for graphed in (False, True):
        m = _model_from(c)
        fopt = trainer.FusedFieldAdam(m, lr=1e-3)
        gs = trainer.GraphedTrainStep(m, fopt)
        losses = []
        for _ in range(6):
            loss, pred = gs.step(b) if graphed else trainer.fused_train_step(m, fopt, b)
            losses.append(loss)
        runs.append((losses, fopt.flat.clone(), fopt.step_count))
    _check_no_watchdog()
    (l0, f0, s0), (l1, f1, s1) = runs
    assert s0 == s1 == 6                      # the device-resident step counter advances inside the replayed graph
    assert l0 == l1
    assert torch.equal(f0, f1)


# ---------------------------------------------------------------------------------------------
# shape coverage of the compiled set: depth, skip position, wide k (many wgrad slices), view branch at max depth
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,L,skip,view_name", [(64, 3, 1, None), (200, 8, 6, None), (2047, 4, 2, None), (96, 7, 3, "view_extrinsic"),
                                                (130, 3, 1, "view_intrinsic")])
def test_shape_coverage_against_oracle(k, L, skip, view_name):
    import types
    from conftest import VIEW_CFGS
    view = VIEW_CFGS[view_name] if view_name else None
    g = torch.Generator().manual_seed(k + L)
    V, F_, B = 400, 700, 517
    E = torch.randn(V, k, generator=g) / 6
    normals = torch.nn.functional.normalize(torch.randn(F_, 3, generator=g), dim=-1)
    fid = torch.randint(0, F_, (B,), generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1)
    dirs = torch.where(((-dirs * normals[fid]).sum(-1) < 0)[:, None], -dirs, dirs)
    vids = torch.randint(0, V, (B, 3), generator=g)
    u = torch.rand(B, 3, generator=g)
    bary = u / u.sum(-1, keepdim=True)
    target = torch.rand(B, 3, generator=g)
    params = O.make_params(k, num_layers=L, skip=skip, view=view, seed=L)
    cfg = dict(k=k, num_layers=L, mlp_hidden_dim=128, skip_layer_idx=skip)
    mesh = None
    kw = dict(num_layers=L, skip=skip)
    if view is not None:
        cfg["view_dependence"] = view
        mesh = types.SimpleNamespace(face_normals=normals.numpy())
        kw.update(view=view, unit_ray_dirs=dirs, hit_face_idxs=fid, face_normals=normals)
    m = inf_b200.make_model(cfg, mesh=mesh)
    m.load_state_dict(params)
    m = m.to(DEV)
    b = {"eigenfunction_table": inf_b200.EigenTable.from_dense(E, device=DEV), "vertex_idxs_of_hit_faces": vids.to(DEV),
         "barycentric_coords": bary.to(DEV), "expected_rgbs": target.to(DEV)}
    if view is not None:
        b["unit_ray_dirs"] = dirs.to(DEV)
        b["hit_face_idxs"] = fid.to(DEV)
    sums, pred, flat = m.l1_step(b)
    _check_no_watchdog()
    loss, pref, grads = O.loss_and_grads(params, E, vids, bary, target, **kw)
    err = (pred.cpu() - pref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN, (err.max().item(), err.mean().item())
    assert abs(sums[0].item() / (3 * B) - loss) <= 1e-3
    _check_grads(m, grads, tol=GRAD_REL_VIEW, cos_min=GRAD_COS_VIEW)
    _, _, ge = O.emulated_fp16_grads(params, E, vids, bary, target, L, skip, vertex_path=(ops.DEFAULT_PATH == "vertex"),
                                     proj=_device_proj(m, b["eigenfunction_table"]),
                                     **{k_: v for k_, v in kw.items() if k_ not in ("num_layers", "skip")})
    # eight layers deep, one-ulp differences of the fp16 activations (fp32 accumulation on the device, exact in the
    # restatement) flip a few masks of their own: measured 5.6e-3 per ray / 1.8e-2 per vertex path on the (200, 8, 6) case
    deep = L > 6
    _check_grads(m, ge, tol=2.5e-2 if deep else GRAD_REL_EMULATED, cos_min=0.9995 if deep else 0.9999)


def test_evaluate_matches_reference_val_loop_and_host_loader_shuffle():
    # Trainer.evaluate (trainer.py:164-187): mean L1 over rays-weighted batches + epoch PSNR = -10 log10(sum sq err / rays)
    from inf_b200 import trainer
    from inf_b200.ray_dataloader import RayDataLoader, HostRayLoader
    c = load_golden("tf_k64")
    m = _model_from(c)
    dl = RayDataLoader(c["table"], "efuncs", c["vids"], c["bary"], c["target"], None, None, batch_size=100, shuffle=False,
                       drop_last=False, device=DEV)
    val_loss, psnr = trainer.evaluate(m, dl)
    ref_pred = c["pred"]
    assert abs(val_loss - (ref_pred - c["target"]).abs().mean().item()) < 1e-3
    assert abs(psnr - O.epoch_psnr(float(((ref_pred - c["target"]) ** 2).sum()), ref_pred.shape[0])) < 0.05
    # shuffled host loader: every ray exactly once per epoch, rows stay paired
    tab = inf_b200.EigenTable.from_dense(c["table"], device=DEV)
    hl = HostRayLoader(tab, c["vids"], c["bary"], c["target"], batch_size=64, shuffle=True, device=DEV,
                       generator=torch.Generator().manual_seed(0))
    got_v, got_t = [], []
    for b in hl:
        got_v.append(b["vertex_idxs_of_hit_faces"].cpu().clone())
        got_t.append(b["expected_rgbs"].cpu().clone())
    gv, gt = torch.cat(got_v), torch.cat(got_t)
    assert gv.shape == c["vids"].shape
    key = lambda v, t: sorted(map(tuple, torch.cat([v.double(), t.double()], 1).tolist()))
    assert key(gv, gt) == key(c["vids"], c["target"])


@pytest.mark.parametrize("name", ["tf_k64", "view_intrinsic"])
def test_fused_batch_selection_equals_index_select_batches(name):
    # ray_dataloader.py:103-145: batch_idxs -> index selects -> model.  With fused_select the kernels read the whole split
    # through the batch's ray indices; results must be bit-identical to the materialised batches, in every mode.
    from inf_b200 import trainer
    from inf_b200.ray_dataloader import RayDataLoader
    c = load_golden(name)
    view = c["view"] is not None
    dirs, fids = (c["unit_ray_dirs"], c["face_idxs"]) if view else (None, None)
    outs = []
    for fused in (False, True):
        gen = torch.Generator(device=DEV).manual_seed(3)
        dl = RayDataLoader(c["table"], "efuncs", c["vids"], c["bary"], c["target"], dirs, fids, batch_size=50, shuffle=True,
                           drop_last=False, device=DEV, generator=gen, fused_select=fused)
        m = _model_from(c)
        fopt = trainer.FusedFieldAdam(m, lr=1e-3)
        gs = trainer.GraphedTrainStep(m, fopt)
        losses, preds = [], []
        for i, b in enumerate(dl):
            if i < 2:
                with torch.no_grad():
                    preds.append(m(b).clone())                 # inference through the selection
            loss, pred = gs.step(b) if i >= 2 else trainer.fused_train_step(m, fopt, b)
            losses.append(loss)
            preds.append(pred.clone())
        m.eval()
        val = trainer.evaluate(m, dl)
        outs.append((losses, torch.cat(preds), fopt.flat.clone(), val))
    _check_no_watchdog()
    (l0, p0, f0, v0), (l1, p1, f1, v1) = outs
    assert l0 == l1 and torch.equal(p0, p1) and torch.equal(f0, f1)
    assert abs(v0[0] - v1[0]) < 1e-6 and abs(v0[1] - v1[1]) < 1e-4


@pytest.mark.parametrize("table_dtype", ["f16", "f32"])
def test_single_chunk_tiles_many_tiles_per_cta(table_dtype):
    # k_pad = 64: one k-chunk per tile, so the gather's load cursor runs two TILES ahead of the consumer (fp16 table);
    # 40 000 rays = 313 tiles > 2 per CTA.  Regression test for the metadata look-ahead of that case.
    torch.manual_seed(5)
    V, k, B = 3000, 64, 40000
    E = torch.randn(V, k) / 4
    params = O.make_params(k, seed=6)
    m = inf_b200.make_model(dict(k=k, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3))
    m.load_state_dict(params)
    m = m.to(DEV)
    vids = torch.randint(0, V, (B, 3))
    u = torch.rand(B, 3)
    bary = u / u.sum(-1, keepdim=True)
    target = torch.rand(B, 3)
    b = {"eigenfunction_table": inf_b200.EigenTable.from_dense(E, device=DEV, dtype=table_dtype),
         "vertex_idxs_of_hit_faces": vids.to(DEV), "barycentric_coords": bary.to(DEV), "expected_rgbs": target.to(DEV)}
    with torch.no_grad():
        p_inf = m(b).cpu()
    sums, pred, flat = m.l1_step(b)
    _check_no_watchdog()
    assert torch.equal(p_inf, pred.cpu())
    idx = torch.arange(0, B, 37)
    ref = O.field_forward(params, E, vids[idx], bary[idx])
    err = (p_inf[idx] - ref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN
    loss_ref = (O.field_forward(params, E, vids, bary) - target).abs().mean().item()
    assert abs(sums[0].item() / (3 * B) - loss_ref) < 1e-3


@pytest.mark.parametrize("scale", [2.0, 3.0])
def test_rgb_parity_at_trained_scale_weights(scale):
    # all-bf16 operands pass at Xavier init but leave the RGB tolerance once training has grown the weights
    # (|W0|_inf 0.07 -> 0.13+); the fp16-operand path must hold it there.  Weights scaled up instead of a long CPU training.
    c = load_golden("tf_k1023")
    params = {n: (t * scale if n.endswith("weight") else t) for n, t in c["params"].items()}
    m = inf_b200.make_model(dict(k=c["k"], num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3))
    m.load_state_dict(params)
    m = m.to(DEV)
    for dt in ("f32", "f16"):
        with torch.no_grad():
            pred = m(_batch(c, dt)).cpu()
        ref = O.field_forward(params, c["table"], c["vids"], c["bary"])
        err = (pred - ref).abs()
        assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN, (scale, dt, err.max().item(), err.mean().item())
    _check_no_watchdog()


def test_captured_graph_survives_workspace_growth():
    # a CUDA graph captured at a small batch keeps pointing into the scratch buffer of that time; a later, larger call
    # outgrows the buffer -- the old one must stay alive or the replay would write through a dangling pointer
    from inf_b200 import trainer
    c = load_golden("tf_k64")
    b = _batch(c)
    m = _model_from(c)
    fopt = trainer.FusedFieldAdam(m, lr=1e-3)
    gs = trainer.GraphedTrainStep(m, fopt)
    l = [gs.step(b)[0] for _ in range(3)]                       # eager, capture, replay
    # force the workspace to grow: a training call 200x larger
    torch.manual_seed(0)
    V, B = c["table"].shape[0], 60000
    big = {"eigenfunction_table": b["eigenfunction_table"], "vertex_idxs_of_hit_faces": torch.randint(0, V, (B, 3), device=DEV),
           "barycentric_coords": torch.full((B, 3), 1.0 / 3.0, device=DEV), "expected_rgbs": torch.rand(B, 3, device=DEV)}
    m2 = _model_from(c)
    m2.l1_step(big)
    filler = [torch.full((1 << 18,), float("nan"), device=DEV) for _ in range(64)]   # reuse freed blocks, if any, with poison
    l += [gs.step(b)[0] for _ in range(3)]
    _check_no_watchdog()
    del filler
    ref_m = _model_from(c)
    ropt = trainer.FusedFieldAdam(ref_m, lr=1e-3)
    lr_ = [trainer.fused_train_step(ref_m, ropt, b)[0] for _ in range(6)]
    assert l == lr_
    assert torch.equal(fopt.flat, ropt.flat)


# ---------------------------------------------------------------------------------------------
# randomized configurations: every knob of the compiled set at once, against the oracle
# ---------------------------------------------------------------------------------------------
def _random_case(seed):
    import random
    rnd = random.Random(seed)
    L = rnd.randint(3, 8)
    use_view = rnd.random() < 0.5
    if use_view:
        L = min(L, 7)
    skip = rnd.randint(1, L - 2)
    k = rnd.choice([1, 7, 63, 64, 65, 100, 129, 300, 513])
    view = None
    if use_view:
        strategy = rnd.choice(["intrinsic", "extrinsic"])
        d = 1 if strategy == "intrinsic" else 3
        inc = rnd.random() < 0.5
        # frequencies above ~2^12 turn the last ulp of acos() into O(1e-3) feature noise in fp32 (in the reference as
        # well), which a 2 % gradient bound cannot absorb: keep the embedding in the well-conditioned range
        emb = rnd.randint(0 if inc else 1, min(10, (64 - (d if inc else 0)) // (2 * d)))
        view = dict(strategy=strategy, in_dim_view_dir=d, include_view_dir=inc, embed_size=emb,
                    bottleneck_vec_dim=rnd.choice([8, 64, 100, 128]), directional_hidden_dim=rnd.choice([16, 48, 64, 128]))
    return dict(L=L, skip=skip, k=k, view=view, B=rnd.choice([1, 5, 127, 128, 129, 300, 700]), V=rnd.randint(3, 400),
                table_dtype=rnd.choice(["f32", "f16"]), idx32=rnd.random() < 0.5, fused_select=rnd.random() < 0.5)


@pytest.mark.parametrize("seed", list(range(96)))
def test_randomized_configurations(seed):
    import types
    cs = _random_case(seed)
    L, skip, k, view, B, V = cs["L"], cs["skip"], cs["k"], cs["view"], cs["B"], cs["V"]
    g = torch.Generator().manual_seed(1000 + seed)
    F_ = 2 * V + 5
    E = torch.randn(V, k, generator=g) / 4
    normals = torch.nn.functional.normalize(torch.randn(F_, 3, generator=g), dim=-1)
    N = B + 37                                               # the split is larger than the batch (fused selection)
    fid = torch.randint(0, F_, (N,), generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
    dirs = torch.where(((-dirs * normals[fid]).sum(-1) < 0)[:, None], -dirs, dirs)
    vids = torch.randint(0, V, (N, 3), generator=g)
    u = torch.rand(N, 3, generator=g)
    bary = u / u.sum(-1, keepdim=True)
    target = torch.rand(N, 3, generator=g)
    sel = torch.randperm(N, generator=g)[:B]
    params = O.make_params(k, num_layers=L, skip=skip, view=view, seed=seed)
    cfg = dict(k=k, num_layers=L, mlp_hidden_dim=128, skip_layer_idx=skip)
    mesh = None
    kw = dict(num_layers=L, skip=skip)
    if view is not None:
        cfg["view_dependence"] = view
        mesh = types.SimpleNamespace(face_normals=normals.numpy())
        kw.update(view=view, unit_ray_dirs=dirs[sel], hit_face_idxs=fid[sel], face_normals=normals)
    m = inf_b200.make_model(cfg, mesh=mesh)
    m.load_state_dict(params)
    m = m.to(DEV)
    it = torch.int32 if cs["idx32"] else torch.int64
    tab = inf_b200.EigenTable.from_dense(E, device=DEV, dtype=cs["table_dtype"])
    if cs["fused_select"]:
        b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": vids.to(DEV).to(it), "barycentric_coords": bary.to(DEV),
             "expected_rgbs": target.to(DEV), "ray_idxs": sel.to(DEV)}
        if view is not None:
            b["unit_ray_dirs"], b["hit_face_idxs"] = dirs.to(DEV), fid.to(DEV)
    else:
        b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": vids[sel].to(DEV).to(it), "barycentric_coords": bary[sel].to(DEV),
             "expected_rgbs": target[sel].to(DEV)}
        if view is not None:
            b["unit_ray_dirs"], b["hit_face_idxs"] = dirs[sel].to(DEV), fid[sel].to(DEV)
    with torch.no_grad():
        p_inf = m(b).cpu()
    sums, pred, flat = m.l1_step(b)
    _check_no_watchdog()
    assert torch.equal(p_inf, pred.cpu()), cs
    loss, pref, grads = O.loss_and_grads(params, E, vids[sel], bary[sel], target[sel], **kw)
    err = (pred.cpu() - pref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN, (cs, err.max().item(), err.mean().item())
    assert abs(sums[0].item() / (3 * B) - loss) <= 1e-3, cs
    ekw = {k_: v for k_, v in kw.items() if k_ not in ("num_layers", "skip")}
    _, _, ge = O.emulated_fp16_grads(params, E, vids[sel], bary[sel], target[sel], L, skip,
                                     table_f16=(cs["table_dtype"] == "f16"), vertex_path=(ops.DEFAULT_PATH == "vertex"),
                                     proj=_device_proj(m, tab), **ekw)
    # tiny batches leave most gradient entries at a few ulps of fp16 noise: compare the whole flat vector
    got = torch.cat([p.grad.detach().cpu().reshape(-1) for p in m.parameters()])
    ref = torch.cat([torch.from_numpy(ge[n]).float().reshape(-1) for n, _ in m.named_parameters()])
    rel = (got - ref).norm().item() / (ref.norm().item() + 1e-20)
    assert rel <= 2e-2, (cs, rel)


# ---------------------------------------------------------------------------------------------
# section 8(f) row 4: UV-baking call site and the preprocessed-ray directory feeding the loaders
# ---------------------------------------------------------------------------------------------
def test_bake_pred_rgbs_matches_reference_texel_path():
    from inf_b200 import bake
    c = load_golden("tf_k100_l5s1")
    m = _model_from(c)
    g = np.random.default_rng(0)
    V = c["table"].shape[0]
    n = 70000                                                      # three 2^15 chunks in the reference, ragged tail
    faces = g.integers(0, V, size=(n, 3))                          # int64 like trimesh faces
    u = g.random((n, 3))
    bary64 = u / u.sum(-1, keepdims=True)                          # float64 like trimesh barycentrics
    feats = c["table"].numpy()
    got = bake.pred_rgbs(m, feats, faces, bary64, device=DEV)
    _check_no_watchdog()
    assert got.dtype == np.float32 and got.shape == (n, 3)
    params = {k_: v for k_, v in c["params"].items()}
    ref = O.bake_pred_rgbs(params, feats, faces, bary64, num_layers=c["num_layers"], skip=c["skip"])
    err = np.abs(got - ref)
    assert err.max() <= RGB_MAX and err.mean() <= RGB_MEAN, (err.max(), err.mean())
    # int32 ids, a different chunking and an empty request give the same texels
    assert np.array_equal(bake.pred_rgbs(m, feats, faces.astype(np.int32), bary64, device=DEV, chunk=4096), got)
    assert bake.pred_rgbs(m, feats, faces[:0], bary64[:0], device=DEV).shape == (0, 3)
    cv = load_golden("view_intrinsic")
    with pytest.raises(NotImplementedError):
        bake.pred_rgbs(_model_from(cv), cv["table"].numpy(), faces[:4] % cv["table"].shape[0], bary64[:4], device=DEV)


def test_ray_store_round_trip_feeds_training_bit_exactly(tmp_path):
    # writer -> five .npy files -> pinned int32 arrays -> HostRayLoader -> fused steps == the same steps from memory
    from inf_b200 import ray_store, trainer
    from inf_b200.ray_dataloader import HostRayLoader
    c = load_golden("view_intrinsic")
    n = c["vids"].shape[0]
    with ray_store.RayCacheWriter(str(tmp_path)) as w:
        for lo in range(0, n, 37):                                  # "views" of 37 hits
            hi = min(n, lo + 37)
            w.append_hits(c["vids"][lo:hi], c["bary"][lo:hi], c["target"][lo:hi], c["unit_ray_dirs"][lo:hi], c["face_idxs"][lo:hi])
    data = ray_store.load_preprocessed_rays(str(tmp_path), pin=True)
    assert data["vertex_idxs_of_hit_faces"].dtype == torch.int32 and data["vertex_idxs_of_hit_faces"].is_pinned()
    assert torch.equal(data["vertex_idxs_of_hit_faces"].long(), c["vids"]) and torch.equal(data["barycentric_coords"], c["bary"])
    tab = inf_b200.EigenTable.from_dense(c["table"], device=DEV)

    def run(loader):
        m = _model_from(c)
        opt = trainer.FusedFieldAdam(m, lr=1e-3)
        losses = []
        for b in loader:
            losses.append(trainer.fused_train_step(m, opt, b)[0])
        return losses, opt.flat.clone()
    la, pa = run(HostRayLoader(tab, data["vertex_idxs_of_hit_faces"], data["barycentric_coords"], data["expected_rgbs"],
                               data["unit_ray_dirs"], data["face_idxs"], batch_size=48, device=DEV))
    lb, pb = run(HostRayLoader(tab, c["vids"], c["bary"], c["target"], c["unit_ray_dirs"], c["face_idxs"], batch_size=48, device=DEV))
    _check_no_watchdog()
    assert la == lb and torch.equal(pa, pb)


@pytest.mark.parametrize("name", ["tf_k64", "view_intrinsic"])
def test_renderer_render_matches_reference_canvas(name):
    # Renderer.render (renderer.py:64-146) with a recorded intersector: pixel rays -> hits -> RGB -> canvas, masked and not
    import types
    from inf_b200.renderer import Renderer
    from inf_b200.ray_store import pixel_rays
    c = load_golden(name)
    m = _model_from(c)
    H, W = 20, 24
    g = torch.Generator().manual_seed(3)
    n = c["vids"].shape[0]
    cam = torch.tensor([[1.0, 0, 0, 0.1], [0, 1, 0, -0.2], [0, 0, 1, -3.0]])
    K = torch.tensor([[30.0, 0, W / 2 - 0.5], [0, 28.0, H / 2 - 0.5], [0, 0, 1]])
    for masked in (False, True):
        mask = (torch.rand(H * W, generator=g) < 0.7) if masked else None
        L = int(mask.sum()) if masked else H * W
        hits = torch.randperm(L, generator=g)[:min(L, n) - 7]              # rays that hit, in intersector order
        sel = torch.randint(0, n, (hits.numel(),), generator=g)            # which golden hit each of them is
        seen = {}

        def intersect(origins, dirs):
            assert origins.shape == dirs.shape == (L, 3)
            seen["dirs"] = dirs
            fid = c["face_idxs"][sel] if c.get("view") is not None else torch.zeros(sel.numel(), dtype=torch.int64)
            return c["vids"][sel], c["bary"][sel], hits, fid
        mesh = types.SimpleNamespace(face_normals=c["face_normals"].numpy()) if c.get("view") is not None else None
        r = Renderer(m, mesh, eigenfunctions=c["table"], device=DEV, H=H, W=W, intersect=intersect)
        img = r.render(cam, K, obj_mask_1d=mask)
        _check_no_watchdog()
        assert img.shape == (H, W, 3) and img.dtype == np.float32
        # restatement of renderer.py:113-146 with the oracle as the model
        kw = dict(num_layers=c["num_layers"], skip=c["skip"])
        if c.get("view") is not None:
            _, dirs = pixel_rays(cam, K, mask if masked else torch.ones(H * W, dtype=torch.bool), H=H, W=W)
            assert torch.equal(dirs, seen["dirs"])
            kw.update(view=c["view"], unit_ray_dirs=dirs[hits], hit_face_idxs=c["face_idxs"][sel], face_normals=c["face_normals"])
        pred = O.field_forward(c["params"], c["table"], c["vids"][sel], c["bary"][sel], **kw)
        canvas = torch.ones(L, 3)
        canvas[hits] = pred
        if masked:
            full = torch.ones(H * W, 3)
            full[mask] = canvas
            canvas = full
        ref = canvas.reshape(H, W, 3).numpy()
        bg = np.ones((H * W, 3), bool)
        idx = torch.nonzero(mask).reshape(-1)[hits] if masked else hits
        bg[idx.numpy()] = False
        assert np.array_equal(img.reshape(-1, 3)[bg[:, 0]], ref.reshape(-1, 3)[bg[:, 0]])       # background untouched: exactly 1.0
        err = np.abs(img - ref)
        assert err.max() <= RGB_MAX and err.mean() <= RGB_MEAN
    r.set_height(10); r.set_width(12)
    assert (r.H, r.W) == (10, 12)
    with pytest.raises(NotImplementedError):
        Renderer(m, None, eigenfunctions=c["table"], feature_strategy="xyz", device=DEV, H=4, W=4)


@pytest.mark.parametrize("name", ["tf_k1023", "view_intrinsic"])
def test_render_frames_pipeline_is_bit_identical_to_render_per_view(name):
    # render_video.py:117-134: a turntable of views.  render_frames overlaps the caller's intersector for frame n+1 with
    # the GPU work of frame n and the D2H of frame n-1; the images must equal render() called view by view, bit for bit.
    import time
    import types
    from inf_b200.renderer import Renderer
    c = load_golden(name)
    m = _model_from(c)
    H, W = 24, 32
    n = c["vids"].shape[0]
    K = torch.tensor([[30.0, 0, W / 2 - 0.5], [0, 28.0, H / 2 - 0.5], [0, 0, 1]])
    log = []

    def intersect(origins, dirs):
        # a recorded caster: which rays hit depends on the view (through the ray directions), deterministic, slow on purpose
        g = torch.Generator().manual_seed(int(dirs[0, 0].abs().item() * 1e6) % 1000)
        time.sleep(0.01)
        L = origins.shape[0]
        hits = torch.randperm(L, generator=g)[:min(L, n) - 5]
        sel = torch.randint(0, n, (hits.numel(),), generator=g)
        fid = c["face_idxs"][sel] if c.get("view") is not None else torch.zeros(sel.numel(), dtype=torch.int64)
        log.append(int(hits[0]))
        return c["vids"][sel], c["bary"][sel], hits, fid
    mesh = types.SimpleNamespace(face_normals=c["face_normals"].numpy()) if c.get("view") is not None else None
    r = Renderer(m, mesh, eigenfunctions=c["table"], device=DEV, H=H, W=W, intersect=intersect)
    views = []
    for a in np.linspace(0, 300, 6):
        st, ct = float(np.sin(np.deg2rad(a))), float(np.cos(np.deg2rad(a)))
        cam = torch.tensor([[ct, -st, 0, 0.1], [st, ct, 0, -0.2], [0, 0, 1, -3.0]])
        mask = (torch.rand(H * W, generator=torch.Generator().manual_seed(int(a))) < 0.8) if int(a) % 120 == 0 else None
        views.append((cam, K, mask))
    frames = list(r.render_frames(iter(views)))
    _check_no_watchdog()
    assert len(frames) == len(views)
    order = list(log)
    singles = [r.render(cam, K_, obj_mask_1d=mask) for cam, K_, mask in views]
    assert log[len(order):] == order                      # the caster saw the same views in the same order
    for f, s_ in zip(frames, singles):
        assert f.shape == (H, W, 3) and f.dtype == np.float32
        assert np.array_equal(f, s_)


def test_trainer_epoch_loop_matches_reference_loop_and_resumes(tmp_path):
    # Trainer (trainer.py:18-340): epoch metrics against the same loop run with the fp32 oracle, best-model / last-epoch /
    # checkpoint files, and a resumed run that ends bit-identical to an uninterrupted one
    from inf_b200 import trainer
    from inf_b200.ray_dataloader import RayDataLoader
    c = load_golden("tf_k64")
    n, ntr, B, lr = c["vids"].shape[0], 192, 64, 1e-3
    tr_sl, va_sl = slice(0, ntr), slice(ntr, n)

    # ---- the reference's loop with the oracle as the model (CPU, fp32) ----
    q = {k_: v.detach().clone().requires_grad_(True) for k_, v in c["params"].items()}
    opt = torch.optim.Adam(list(q.values()), lr=lr)
    ref = []
    for epoch in range(3):
        acc_l = acc_sq = 0.0
        for lo in range(0, ntr, B):
            sl = slice(lo, lo + B)
            pred = O.field_forward(q, c["table"], c["vids"][sl], c["bary"][sl])
            loss = O.l1_loss(pred, c["target"][sl])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            acc_l += loss.item() * B
            acc_sq += float(((pred.detach() - c["target"][sl]) ** 2).sum())
        with torch.no_grad():
            pv = O.field_forward(q, c["table"], c["vids"][va_sl], c["bary"][va_sl])
        ref.append((acc_l / ntr, O.epoch_psnr(acc_sq, ntr), float(O.l1_loss(pv, c["target"][va_sl])),
                    O.epoch_psnr(float(((pv - c["target"][va_sl]) ** 2).sum()), n - ntr)))

    def make(out_dir, epochs, fused_select):
        m = _model_from(c)
        tab = inf_b200.EigenTable.from_dense(c["table"], device=DEV)
        mk = lambda sl, fs: RayDataLoader(tab, "efuncs", c["vids"][sl], c["bary"][sl], c["target"][sl], None, None, B, shuffle=False,
                                          drop_last=False, device=DEV, fused_select=fs)
        cfg = {"training": {"out_dir": str(out_dir), "epochs": epochs, "print_every": 100, "render_every": 100,
                            "checkpoint_every": 1, "lr": lr, "loss_type": "L1"},
               "data": {"img_height": 8, "img_width": 8}, "model": {}}
        fopt = trainer.FusedFieldAdam(m, lr=lr)
        data = {"train": mk(tr_sl, fused_select), "val": mk(va_sl, False), "test": mk(va_sl, False)}
        return trainer.Trainer(m, fopt, trainer.get_loss_fn("L1"), None, data, None, cfg, DEV, writer=trainer._ScalarLog())

    t3 = make(tmp_path / "a", 3, True)
    t3.train()
    _check_no_watchdog()
    sc = t3.writer.scalars
    for e, (trl, trp, val, vap) in enumerate(ref):
        assert abs(sc["Train_Loss"][e][1] - trl) < 2e-3 and abs(sc["Val_Loss"][e][1] - val) < 2e-3, (e, sc["Train_Loss"][e], trl)
        assert abs(sc["Train Epoch-PSNR"][e][1] - trp) < 0.1 and abs(sc["Val Epoch-PSNR"][e][1] - vap) < 0.1
    assert [s for s, _ in sc["Val_Loss"]] == [0, 1, 2] and len(sc["Test Loss"]) == 1
    for f in ("model.pt", "model_last_epoch.pt", "checkpoint.pt"):
        assert (tmp_path / "a" / f).exists(), f
    best = torch.load(tmp_path / "a" / "model.pt")
    assert list(best.keys()) == c["param_names"]
    ck = torch.load(tmp_path / "a" / "checkpoint.pt", weights_only=False)
    assert ck["epoch"] == 2 and set(ck) >= {"model_state_dict", "optimizer_state_dict", "pytorch_random_state"}

    # ---- resume: 3 epochs + restart to 5 == 5 epochs straight ----
    t5 = make(tmp_path / "b", 5, False)
    t5.train()
    t_res = make(tmp_path / "a", 5, False)                   # finds a/checkpoint.pt (epoch 2) and continues at epoch 3
    t_res.train()
    assert [s for s, _ in t_res.writer.scalars["Val_Loss"]] == [3, 4]
    a, b = t5.model.state_dict(), t_res.model.state_dict()
    for k_ in a:
        assert torch.equal(a[k_], b[k_]), k_
    assert t5.optim.step_count == t_res.optim.step_count == 15


def test_get_data_builds_the_loaders_of_a_config(tmp_path):
    # config.py:58-103 + ray_dataloader.py:7-58: stored eigen-decomposition + preprocessed-ray directories -> loaders
    from inf_b200 import ray_store
    from inf_b200.ray_dataloader import get_data
    c = load_golden("view_intrinsic")
    n, k = c["vids"].shape[0], c["table"].shape[1]
    g = np.random.default_rng(0)
    stored = g.standard_normal((c["table"].shape[0], k + 5))                      # float64, more columns than the model uses
    np.save(tmp_path / "efuncs.npy", stored)
    for split, sl in (("train", slice(0, 100)), ("eval", slice(100, n))):
        with ray_store.RayCacheWriter(str(tmp_path / split)) as w:
            w.append_hits(c["vids"][sl], c["bary"][sl], c["target"][sl], c["unit_ray_dirs"][sl], c["face_idxs"][sl])
    cfg = {"data": {"eigenfunctions_path": str(tmp_path / "efuncs.npy"), "preproc_data_path_train": str(tmp_path / "train"),
                    "preproc_data_path_eval": str(tmp_path / "eval")},
           "model": {"k": k}, "training": {"batch_size": 32}}
    data = get_data(cfg, DEV)
    assert set(data) == {"train", "val"} and data["train"].table is data["val"].table
    want = stored[:, :k] / (stored[:, :k].max(0, keepdims=True) - stored[:, :k].min(0, keepdims=True))
    assert torch.equal(data["train"].table.data[:, :k].cpu(), torch.from_numpy(want).float())        # 'standard' rescale, bit-exact
    assert len(data["train"]) == 3 and len(data["val"]) == (n - 100 + 31) // 32                      # drop_last True / False
    seen = 0
    for b in data["val"]:                                                      # unshuffled: rows in stored order, values untouched
        m = b["expected_rgbs"].shape[0]
        sl = slice(100 + seen, 100 + seen + m)
        assert b["vertex_idxs_of_hit_faces"].dtype == torch.int32
        assert torch.equal(b["vertex_idxs_of_hit_faces"].cpu().long(), c["vids"][sl]) and torch.equal(b["barycentric_coords"].cpu(), c["bary"][sl])
        assert torch.equal(b["unit_ray_dirs"].cpu(), c["unit_ray_dirs"][sl]) and torch.equal(b["hit_face_idxs"].cpu().long(), c["face_idxs"][sl])
        seen += m
    assert seen == n - 100
    b = next(iter(data["train"]))                                              # shuffled, selection folded into the kernels
    assert b["ray_idxs"].shape == (32,) and b["vertex_idxs_of_hit_faces"].shape == (100, 3)


def test_trainer_with_torch_adam_and_l2_loss_uses_the_autograd_path(tmp_path):
    # loss_type L2 (config.py:115) + torch.optim.Adam + ReduceLROnPlateau: the generic branch of Trainer._train_step
    from inf_b200 import trainer
    from inf_b200.ray_dataloader import RayDataLoader
    c = load_golden("tf_k64")
    n, ntr, B, lr = c["vids"].shape[0], 192, 64, 1e-3
    q = {k_: v.detach().clone().requires_grad_(True) for k_, v in c["params"].items()}
    opt = torch.optim.Adam(list(q.values()), lr=lr)
    acc = 0.0
    for lo in range(0, ntr, B):
        sl = slice(lo, lo + B)
        loss = torch.nn.functional.mse_loss(O.field_forward(q, c["table"], c["vids"][sl], c["bary"][sl]), c["target"][sl])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        acc += loss.item() * B
    m = _model_from(c)
    tab = inf_b200.EigenTable.from_dense(c["table"], device=DEV)
    mk = lambda sl: RayDataLoader(tab, "efuncs", c["vids"][sl], c["bary"][sl], c["target"][sl], None, None, B, shuffle=False,
                                  drop_last=False, device=DEV)
    cfg = {"training": {"out_dir": str(tmp_path), "epochs": 1, "use_lr_scheduler": True}, "data": {}, "model": {}}
    t = trainer.Trainer(m, torch.optim.Adam(m.parameters(), lr=lr), trainer.get_loss_fn("L2"), None,
                        {"train": mk(slice(0, ntr)), "val": mk(slice(ntr, n))}, None, cfg, DEV, writer=trainer._ScalarLog())
    assert t._graphed is None and t.lr_scheduler is not None
    loss, pred = t._train_step(next(iter(mk(slice(0, B)))))            # trainer.py:71-84 contract
    assert isinstance(loss, float) and pred.shape == (B, 3)
    m.load_state_dict(c["params"])                                      # undo that step
    t.optim = torch.optim.Adam(m.parameters(), lr=lr)
    t.lr_scheduler = None
    t.train()
    _check_no_watchdog()
    assert abs(t.writer.scalars["Train_Loss"][0][1] - acc / ntr) < 1e-3
    with pytest.raises(NotImplementedError):
        trainer.Trainer(m, trainer.FusedFieldAdam(m), trainer.get_loss_fn("L1"), None, {"train": [], "val": []}, None, cfg, DEV)
