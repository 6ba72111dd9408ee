import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["tf_k64", "tf_k1023", "tf_k100_l5s1", "view_intrinsic", "view_extrinsic"]
VIEW_CFGS = {
    "view_intrinsic": dict(strategy="intrinsic", in_dim_view_dir=1, include_view_dir=True, embed_size=4,
                           bottleneck_vec_dim=128, directional_hidden_dim=64),
    "view_extrinsic": dict(strategy="extrinsic", in_dim_view_dir=3, include_view_dir=True, embed_size=4,
                           bottleneck_vec_dim=128, directional_hidden_dim=64),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Golden vectors produced by the live reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    names = [str(n) for n in g["param_names"]] if "param_names" in g else []
    case = {
        "name": name,
        "table": torch.from_numpy(g["table"]),
        "vids": torch.from_numpy(g["vids"]),
        "bary": torch.from_numpy(g["bary"]),
        "x": torch.from_numpy(g["x"]),
    }
    if names:
        case.update(
            target=torch.from_numpy(g["target"]),
            face_idxs=torch.from_numpy(g["face_idxs"]),
            face_normals=torch.from_numpy(g["face_normals"]),
            unit_ray_dirs=torch.from_numpy(g["unit_ray_dirs"]),
            pred=torch.from_numpy(g["pred"]),
            loss=float(g["loss"]),
            num_layers=int(g["num_layers"]), skip=int(g["skip"]), k=int(g["k"]),
            params={n: torch.from_numpy(g["param/" + n]) for n in names},
            grads={n: torch.from_numpy(g["grad/" + n]) for n in names},
            view=VIEW_CFGS.get(name),
            param_names=names,
        )
    return case


def oracle_kwargs(case):
    kw = dict(num_layers=case["num_layers"], skip=case["skip"])
    if case["view"] is not None:
        kw.update(view=case["view"], unit_ray_dirs=case["unit_ray_dirs"], hit_face_idxs=case["face_idxs"],
                  face_normals=case["face_normals"])
    return kw


@pytest.fixture(params=GOLDEN_CASES)
def golden_case(request):
    return load_golden(request.param)
