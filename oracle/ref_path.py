"""The reference's own CPU implementation of the hot path, imported from oracle/_ref (staged by oracle/make_ref.py).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used by `bench.py --impl reference` and bench.py's cpu_baseline leg: the
reference's `get_k_eigenfunc_vec_vals` (mesh.py:313-324), `make_model` (model.py:199-258), F.l1_loss (config.py:117-118),
torch.optim.Adam (config.py:108) -- stock code, run on the host cores.  The only restated lines are the ten of
`Trainer._train_step` (trainer.py:71-84): trainer.py itself cannot be imported (tensorboardX, imageio and
ReduceLROnPlateau(verbose=True), a TypeError on this torch)."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "model.py")) and os.path.exists(os.path.join(REF_DIR, "mesh.py"))


def load():
    """(reference model module, reference mesh module); third-party imports that are not on the path become stubs."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    for name in ("igl", "trimesh", "torchinfo", "imageio", "imageio.plugins", "imageio.plugins.freeimage"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["torchinfo"], "summary"):
        sys.modules["torchinfo"].summary = lambda *a, **k: None
    im = sys.modules["imageio"]
    if not hasattr(im, "plugins"):
        im.plugins = sys.modules["imageio.plugins"]
        im.plugins.freeimage = sys.modules["imageio.plugins.freeimage"]
        im.plugins.freeimage.download = lambda *a, **k: None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    ref_model = importlib.import_module("model")
    ref_mesh = importlib.import_module("mesh")
    if not os.path.samefile(os.path.dirname(ref_model.__file__), REF_DIR):
        raise RuntimeError(f"'model' resolved to {ref_model.__file__}, not the staged reference")
    return ref_model, ref_mesh


class ReferenceStep:
    """trainer.py:71-84 on the reference's own modules (CPU, fp32)."""

    def __init__(self, model_config, table, lr, train=True, mesh=None):
        import torch
        self.torch = torch
        self.ref_model, self.ref_mesh = load()
        self.model = self.ref_model.make_model(model_config, mesh=mesh)
        self.table = table
        self.train = train
        self.optim = torch.optim.Adam(self.model.parameters(), lr=lr) if train else None      # config.py:108

    def step(self, vids, bary, target=None):
        torch = self.torch
        batch = {"eigenfunctions": self.ref_mesh.get_k_eigenfunc_vec_vals(self.table, vids, bary)}   # ray_dataloader.py:126-131
        if not self.train:
            with torch.no_grad():
                return self.model(batch)
        batch["expected_rgbs"] = target
        self.model.train()
        pred = self.model(batch)                                   # trainer.py:75
        loss = torch.nn.functional.l1_loss(pred, target)           # trainer.py:76, config.py:117-118
        self.optim.zero_grad(set_to_none=True)                     # trainer.py:78
        loss.backward()                                            # trainer.py:81
        self.optim.step()                                          # trainer.py:82
        return loss.item(), pred                                   # trainer.py:84
