"""Fixed-step training protocols shared by the GPU parity tests and tests/golden/make_psnr_ensemble.py (the reference
side).  Everything is generated from seeds with torch CPU generators, so both sides see identical problems and identical
batch index lists (CPU and CUDA randperm streams differ: never rely on the seed alone)."""
import torch

from oracle import field_oracle as O

PROTOCOLS = {
    # the round-1 protocol: k=64, 400 steps at the shipped lr
    "small_none": dict(view=None, k=64, L=6, skip=3, B=512, steps=400, lr=1e-4, V=1500, F=2500, N=8192, ntrain=4096),
    "small_view": dict(view="view_intrinsic", k=64, L=6, skip=3, B=512, steps=400, lr=1e-4, V=1500, F=2500, N=8192, ntrain=4096),
    # configs[1] at its shipped sizes (intrinsic_cat.yaml: k=1023, batch 4096, lr 1e-4), 2000 steps
    "cat_2000": dict(view=None, k=1023, L=6, skip=3, B=4096, steps=2000, lr=1e-4, V=50176, F=100352, N=65536 + 16384,
                     ntrain=65536, members=3),
    "view_2000": dict(view="view_intrinsic", k=1023, L=6, skip=3, B=4096, steps=2000, lr=1e-4, V=50176, F=100352,
                      N=65536 + 16384, ntrain=65536, members=5),
}


def problem(pr, seed=0):
    """(E, vids, bary, target, dirs, fid, normals): a smooth procedural colour of the hit point (+ a view-dependent tint)."""
    V, F_, k, N = pr["V"], pr["F"], pr["k"], pr["N"]
    g = torch.Generator().manual_seed(seed)
    E = torch.randn(V, k, generator=g)
    E = (E / (E.max(0, keepdim=True).values - E.min(0, keepdim=True).values)).float()     # 'standard' rescale, mesh.py:99-102
    if F_ <= 4096:
        faces = torch.stack([torch.randperm(V, generator=g)[:3] for _ in range(F_)])
    else:
        faces = torch.randint(0, V, (F_, 3), generator=g)
    fid = torch.randint(0, F_, (N,), generator=g)
    vids = faces[fid]
    u = torch.rand(N, 3, generator=g)
    bary = (u / u.sum(-1, keepdim=True)).float()
    normals = torch.nn.functional.normalize(torch.randn(F_, 3, generator=g), dim=-1)
    dirs = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
    dirs = torch.where(((-dirs * normals[fid]).sum(-1) < 0)[:, None], -dirs, dirs).float()
    x = O.gather_interpolate(E, vids, bary)
    mix = torch.randn(k, 3, generator=g) * (1.5 if k <= 64 else 1.5 * (64.0 / k) ** 0.5 * 4)
    logits = x @ mix
    if pr["view"] is not None:
        logits = logits + 0.5 * (-dirs * normals[fid]).sum(-1, keepdim=True)
    target = torch.sigmoid(logits)
    return E, vids, bary, target, dirs, fid, normals


def index_lists(pr):
    gi = torch.Generator().manual_seed(9)
    return [torch.randperm(pr["ntrain"], generator=gi)[:pr["B"]] for _ in range(pr["steps"])]
