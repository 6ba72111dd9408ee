"""BASELINE.json configs at their full sizes on the GPU: size-independent properties (per-ray independence,
linearity of the gradients, chunk invariance) plus the CPU oracle on subsamples.  Tolerances are BASELINE.json's."""
import types

import numpy as np
import pytest
import torch

from conftest import VIEW_CFGS
from oracle import field_oracle as O

import inf_b200
from inf_b200 import ops

pytestmark = pytest.mark.gpu

RGB_MAX, RGB_MEAN = 2e-2, 1e-3
DEV = "cuda:0"


@pytest.fixture(autouse=True, params=["ray", "vertex"])
def eval_path(request):
    """Every config runs on both evaluations of the two K = k contractions (InfDesc.path): per ray and per vertex."""
    old = ops.DEFAULT_PATH
    ops.DEFAULT_PATH = request.param
    yield request.param
    ops.DEFAULT_PATH = old


def _check_no_watchdog():
    torch.cuda.synchronize()
    assert inf_b200.kernel_error_word(DEV) == 0


def icosphere(subdiv):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdiv):
        edge = {}
        verts = list(v)

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in edge:
                m = (verts[a] + verts[b]) / 2.0
                verts.append(m / np.linalg.norm(m))
                edge[key] = len(verts) - 1
            return edge[key]
        nf = []
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.array(verts), np.array(nf, dtype=np.int64)
    return v, f


def cotan_eigenfunctions(v, f, k):
    """First k non-constant LBO eigenfunctions: cotan stiffness + lumped (Voronoi-style) mass and
    eigsh(L, k+1, M, sigma=0, which="LM") as mesh.py:120-168 (robust_laplacian is not installed here)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    n = v.shape[0]
    I, J, W = [], [], []
    for i in range(3):
        a, b, c = f[:, i], f[:, (i + 1) % 3], f[:, (i + 2) % 3]
        e1, e2 = v[b] - v[a], v[c] - v[a]
        cot = (e1 * e2).sum(1) / np.linalg.norm(np.cross(e1, e2), axis=1)
        I += [b, c]; J += [c, b]; W += [0.5 * cot, 0.5 * cot]
    tri_area = 0.5 * np.linalg.norm(np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]), axis=1)
    area = np.zeros(n)
    for i in range(3):
        np.add.at(area, f[:, i], tri_area / 3.0)
    Wm = sp.coo_matrix((np.concatenate(W), (np.concatenate(I), np.concatenate(J))), shape=(n, n)).tocsr()
    L = sp.diags(np.asarray(Wm.sum(1)).ravel()) - Wm
    M = sp.diags(area)
    _, vecs = spl.eigsh(L + 1e-10 * sp.identity(n), k + 1, M, sigma=0, which="LM")
    return vecs[:, 1:]                                               # drop the constant eigenfunction (mesh.py:161-166)


def test_config1_icosphere_64_lbo_eigenfunctions_64k_rays_forward():
    # configs[0]: ~10k-vertex icosphere, first 64 LBO eigenfunctions, 64k rays, forward in 2^15-ray chunks
    v, f = icosphere(5)
    assert v.shape[0] == 10242 and f.shape[0] == 20480
    E = torch.from_numpy(O.rescale_standard(cotan_eigenfunctions(v, f, 64))).float()      # 'standard' rescale, mesh.py:99-102
    g = torch.Generator().manual_seed(0)
    B = 65536
    fid = torch.randint(0, f.shape[0], (B,), generator=g)
    vids = torch.from_numpy(f)[fid]
    u = torch.rand(B, 3, generator=g)
    bary = (u / u.sum(-1, keepdim=True)).float()
    params = O.make_params(64, seed=0)
    m = inf_b200.make_model(dict(k=64, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3))
    m.load_state_dict(params)
    m = m.to(DEV)
    from inf_b200.renderer import FieldRenderer
    r = FieldRenderer(m, E, H=256, W=256, device=DEV, chunk=1 << 15)                      # renderer.py:114
    pred = r.predict_hits(vids, bary).cpu()
    _check_no_watchdog()
    ref = O.field_forward(params, E, vids, bary)                                          # the reference's CPU path, all rays
    err = (pred - ref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN, (err.max().item(), err.mean().item())
    img = r.render_hits(vids, bary, torch.arange(B))
    assert img.shape == (256, 256, 3) and torch.equal(torch.from_numpy(img).reshape(-1, 3), pred)


def test_config3_view_dependent_large_batch_properties():
    # configs[2]: view-direction branch at 131 072 rays per GPU: per-ray independence, gradient linearity, oracle subsample
    view = VIEW_CFGS["view_intrinsic"]
    torch.manual_seed(1)
    V, F_, k, B = 20000, 40000, 1023, 131072
    E = torch.randn(V, k) / 8
    normals = torch.nn.functional.normalize(torch.randn(F_, 3), dim=-1)
    faces = torch.randint(0, V, (F_, 3))
    fid = torch.randint(0, F_, (B,))
    dirs = torch.nn.functional.normalize(torch.randn(B, 3), dim=-1)
    dirs = torch.where(((-dirs * normals[fid]).sum(-1) < 0)[:, None], -dirs, dirs)
    u = torch.rand(B, 3)
    bary = u / u.sum(-1, keepdim=True)
    tgt = torch.rand(B, 3)
    params = O.make_params(k, view=view, seed=2)
    m = inf_b200.make_model(dict(k=k, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3, view_dependence=view),
                            mesh=types.SimpleNamespace(face_normals=normals.numpy()))
    m.load_state_dict(params)
    m = m.to(DEV)
    tab = inf_b200.EigenTable.from_dense(E, device=DEV)
    dv, db, dt, dd, df = (t.to(DEV) for t in (faces[fid], bary, tgt, dirs, fid))

    def run(sl):
        b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": dv[sl].contiguous(), "barycentric_coords": db[sl].contiguous(),
             "expected_rgbs": dt[sl].contiguous(), "unit_ray_dirs": dd[sl].contiguous(), "hit_face_idxs": df[sl].contiguous()}
        sums, pred, flat = m.l1_step(b, global_rays=B)
        return sums.clone(), pred.clone(), flat.clone()
    s_all, p_all, f_all = run(slice(0, B))
    s_a, p_a, f_a = run(slice(0, B // 2))
    s_b, p_b, f_b = run(slice(B // 2, B))
    _check_no_watchdog()
    assert torch.equal(p_all, torch.cat([p_a, p_b]))
    assert torch.allclose(s_all, s_a + s_b, rtol=1e-4)
    assert (f_all - (f_a + f_b)).norm().item() / f_all.norm().item() < 1e-4
    idx = torch.arange(0, B, 257)
    ref = O.field_forward(params, E, faces[fid][idx], bary[idx], view=view, unit_ray_dirs=dirs[idx], hit_face_idxs=fid[idx],
                          face_normals=normals)
    err = (p_all.cpu()[idx] - ref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN


def test_config4_400k_row_table_gather():
    # configs[3]: 400k-point table (1.64 GB fp32, far larger than L2), uniformly random rows = worst-case locality
    torch.manual_seed(2)
    V, k, B = 400000, 1023, 32768
    g = torch.Generator(device=DEV).manual_seed(0)
    E_dev = torch.randn(V, k, device=DEV, generator=g) / 8
    tab = inf_b200.EigenTable.from_dense(E_dev, device=DEV)
    vids = torch.randint(0, V, (B, 3))
    vids[0] = torch.tensor([0, V - 1, V // 2])                      # extreme row ids
    vids[1] = torch.tensor([V - 1, V - 1, V - 1])                   # duplicate ids
    u = torch.rand(B, 3)
    bary = u / u.sum(-1, keepdim=True)
    params = O.make_params(k, seed=4)
    m = inf_b200.make_model(dict(k=k, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3))
    m.load_state_dict(params)
    m = m.to(DEV)
    b = {"eigenfunction_table": tab, "vertex_idxs_of_hit_faces": vids.to(DEV), "barycentric_coords": bary.to(DEV)}
    with torch.no_grad():
        pred = m(b).cpu()
    x = ops.gather_interpolate_dense(tab, vids.to(DEV), bary.to(DEV)).cpu()     # mesh.py:313-324 as one kernel
    _check_no_watchdog()
    idx = torch.cat([torch.arange(0, 8), torch.arange(8, B, 61)])
    rows = E_dev[vids[idx].reshape(-1).to(DEV)].cpu().reshape(len(idx), 3, k)
    x_ref = torch.bmm(bary[idx].unsqueeze(1), rows).squeeze(1)
    assert torch.allclose(x[idx], x_ref, atol=1e-6, rtol=1e-5)
    ref = O.texture_field_forward(params, x_ref, 6, 3)
    err = (pred[idx] - ref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN


def test_config5_1080p_frame_inference_chunk_invariance():
    # configs[4]: one 1080p frame (2 073 600 rays) through the host->host renderer path; chunking must not change a bit
    from inf_b200.renderer import FieldRenderer
    torch.manual_seed(3)
    V, k, M = 50176, 1023, 1920 * 1080
    g = torch.Generator(device=DEV).manual_seed(1)
    tab = inf_b200.EigenTable.from_dense(torch.randn(V, k, device=DEV, generator=g) / 8, device=DEV)
    m = inf_b200.make_model(dict(k=k, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3)).to(DEV)
    vids = torch.randint(0, V, (M, 3), dtype=torch.int32).pin_memory()
    u = torch.rand(M, 3)
    bary = (u / u.sum(-1, keepdim=True)).float().pin_memory()
    r = FieldRenderer(m, tab, H=1080, W=1920, device=DEV)
    a = r.predict_hits_host(vids, bary, chunk=1 << 17).clone()
    b = r.predict_hits_host(vids, bary, chunk=1 << 20).clone()
    _check_no_watchdog()
    assert torch.equal(a, b)
    assert bool(((a > 0) & (a < 1)).all())                                 # sigmoid range (model.py:94)
    idx = torch.arange(0, M, 4099)
    tab_cpu = tab.data[:, :k].cpu()
    ref = O.field_forward({n: p.detach().cpu() for n, p in m.state_dict().items()}, tab_cpu, vids[idx].long(), bary[idx])
    err = (a[idx] - ref).abs()
    assert err.max().item() <= RGB_MAX and err.mean().item() <= RGB_MEAN


def test_config2_cat_training_steps_at_shipped_batch_match_oracle():
    # configs[1] as the reference ships it (intrinsic_cat.yaml:25-34): V ~ 50 k vertices, k = 1023, batch_size 4096, L1, Adam 1e-4.
    # 12 optimisation steps through the whole fast path (int32 ids, batch selection folded into the kernels, fused Adam,
    # CUDA-graph replay) against the fp32 oracle on identical batches: loss trajectory and held-out PSNR.
    from inf_b200 import trainer
    from inf_b200.ray_dataloader import RayDataLoader
    G, k, B, steps, lr = 224, 1023, 4096, 12, 1e-4
    V = G * G
    g = torch.Generator().manual_seed(11)
    E = torch.randn(V, k, generator=g)
    E = (E / (E.max(0, keepdim=True).values - E.min(0, keepdim=True).values)).float()     # 'standard' rescale, mesh.py:99-102
    ii, jj = torch.meshgrid(torch.arange(G), torch.arange(G), indexing="ij")
    v00, v01, v10 = (ii * G + jj).reshape(-1), (ii * G + (jj + 1) % G).reshape(-1), (((ii + 1) % G) * G + jj).reshape(-1)
    v11 = (((ii + 1) % G) * G + (jj + 1) % G).reshape(-1)
    faces = torch.cat([torch.stack([v00, v01, v10], 1), torch.stack([v01, v11, v10], 1)])  # 100 352 triangles of a grid torus
    N = B * 6 + 2048
    fid = torch.randint(0, faces.shape[0], (N,), generator=g)
    vids = faces[fid]
    u = torch.rand(N, 3, generator=g)
    bary = (u / u.sum(-1, keepdim=True)).float()
    # a smooth procedural colour of the hit position (PSNR runs)
    uv = (torch.stack([(vids // G).float(), (vids % G).float()], -1) * bary[..., None]).sum(1) / G
    target = torch.stack([0.5 + 0.4 * torch.sin(6.28 * uv[:, 0]), 0.5 + 0.4 * torch.cos(6.28 * uv[:, 1]),
                          0.5 + 0.3 * torch.sin(6.28 * (uv[:, 0] + uv[:, 1]))], -1).float()
    ntrain = B * 6
    idx_lists = [torch.randperm(ntrain, generator=g)[:B] for _ in range(steps)]
    params = O.make_params(k, seed=21)
    ref_params, ref_losses = O.adam_train_steps(params, E, vids, bary, target, idx_lists, lr=lr)
    hold = slice(ntrain, None)
    with torch.no_grad():
        pr = O.field_forward(ref_params, E, vids[hold], bary[hold])
    psnr_ref = O.epoch_psnr(float(((pr - target[hold]) ** 2).sum()), pr.shape[0])

    m = inf_b200.make_model(dict(k=list(range(k)), num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3))
    m.load_state_dict(params)
    m = m.to(DEV)
    loader = RayDataLoader(E, "efuncs", vids[:ntrain].to(torch.int32), bary[:ntrain], target[:ntrain], None, None, B,
                           shuffle=False, drop_last=True, device=DEV, fused_select=True)
    fopt = trainer.FusedFieldAdam(m, lr=lr)
    gs = trainer.GraphedTrainStep(m, fopt)
    losses = []
    loader.num_batches = steps                       # the explicit order below holds `steps` batches
    loader.N = steps * B
    it = iter(loader)
    loader.set_epoch_order(torch.cat(idx_lists))
    for batch in it:
        loss, _ = gs.step(batch)
        losses.append(loss)
    _check_no_watchdog()
    assert len(losses) == steps
    m.eval()
    with torch.no_grad():
        pg = m({"eigenfunction_table": loader.table, "vertex_idxs_of_hit_faces": vids[hold].to(DEV),
                "barycentric_coords": bary[hold].to(DEV)}).cpu()
    psnr_gpu = O.epoch_psnr(float(((pg - target[hold]) ** 2).sum()), pg.shape[0])
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 2e-3, (losses[-3:], ref_losses[-3:])
    assert abs(psnr_gpu - psnr_ref) <= 0.1, (psnr_gpu, psnr_ref)                 # BASELINE.json tolerance
