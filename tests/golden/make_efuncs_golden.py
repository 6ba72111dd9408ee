"""Golden tables from the LIVE reference ``load_first_k_eigenfunctions`` (mesh.py:53-107).

Run in the build container only:   python tests/golden/make_efuncs_golden.py
A synthetic eigen-decomposition (float64 columns, positive eigenvalues with a numerically-zero negative first one) is
written to a temp directory, the reference function is run unmodified for every (k, embed, rescale) case, and inputs +
outputs go to ``efuncs.npz``.
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, _import_reference  # noqa: E402

CASES = {
    "first_k_standard": dict(k=24, rescale_strategy="standard"),
    "list_k_standard": dict(k=[0, 1, 2, 7, 8, 30, 39], rescale_strategy="standard"),
    "first_k_unscaled": dict(k=40, rescale_strategy="unscaled"),
    "first_k_one_norm": dict(k=16, rescale_strategy="one-norm"),
    "gps": dict(k=20, rescale_strategy="standard", embed_strategy="gps", eigenvalues=True),
    "hks_standard": dict(k=20, rescale_strategy="standard", embed_strategy="hks", eigenvalues=True),
}


def main():
    _, ref_mesh = _import_reference()
    rng = np.random.default_rng(7)
    efuncs = rng.standard_normal((60, 40))
    evals = np.sort(rng.random(40)) * 50 + 0.1
    evals[0] = -3e-12
    store = {"efuncs": efuncs, "evals": evals}
    with tempfile.TemporaryDirectory() as d:
        ep, vp = os.path.join(d, "efuncs.npy"), os.path.join(d, "evals.npy")
        np.save(ep, efuncs)
        np.save(vp, evals)
        for name, c in CASES.items():
            kw = dict(rescale_strategy=c["rescale_strategy"], embed_strategy=c.get("embed_strategy"),
                      eigenvalues_path=vp if c.get("eigenvalues") else None)
            out = ref_mesh.load_first_k_eigenfunctions(ep, c["k"], **kw)
            store["out_" + name] = out.numpy() if hasattr(out, "numpy") else np.asarray(out)
    np.savez_compressed(os.path.join(OUT, "efuncs.npz"), **store)
    print({k: (v.shape, v.dtype) for k, v in store.items()})


if __name__ == "__main__":
    main()
