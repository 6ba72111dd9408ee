"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol declared in
include/inf_field.h, and the host-only size/layout queries agree with the reference's state_dict."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden

import inf_b200
from inf_b200 import _lib, ops


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "inf_field.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(inf_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names
    assert lib.inf_version() == _lib.ABI_VERSION


def _cfg_of(c):
    v = c["view"]
    if v is None:
        return ops.FieldConfig(k=c["k"], num_layers=c["num_layers"], skip=c["skip"])
    return ops.FieldConfig(k=c["k"], num_layers=c["num_layers"], skip=c["skip"], view=v["strategy"],
                           view_embed=v["embed_size"], view_include=v["include_view_dir"],
                           view_bneck=v["bottleneck_vec_dim"], view_hidden=v["directional_hidden_dim"])


@pytest.mark.parametrize("name", ["tf_k64", "tf_k1023", "tf_k100_l5s1", "view_intrinsic", "view_extrinsic"])
def test_param_layout_matches_reference_state_dict(name):
    c = load_golden(name)
    cfg = _cfg_of(c)
    layout = ops.param_layout(cfg)
    assert len(layout) == len(c["param_names"])
    off = 0
    for (o, r, cc), n in zip(layout, c["param_names"]):
        shp = tuple(c["params"][n].shape)
        assert o == off, n
        assert r * cc == c["params"][n].numel(), n
        assert r == shp[0], n
        off += r * cc
    assert ops.num_params(cfg) == off


def test_production_config_sizes():
    cfg = ops.FieldConfig(k=1023, num_layers=6, skip=3)
    assert cfg.k_pad == 1024
    assert ops.num_params(cfg) == 328579
    assert ops.packed_bytes(cfg) == (16 + 4) * 32768 + (5 * 128 + 3 * 128 + 4) * 4
    assert ops.workspace_bytes(cfg, 4096, False) == 65536
    assert ops.workspace_bytes(cfg, 4096, True) > 32 * 9 * 32768


def test_exchange_region_and_projection_sizes():
    import ctypes as C
    cfg = ops.FieldConfig(k=1023, num_layers=6, skip=3)
    lib = _lib.load()
    # 2 parities x (8 source ranks + 1 result slot) x one slot of (value, step number) pairs for total + 2 elements, padded
    # to 32 elements, + the 256-byte control block (csrc/common.cuh: comm_region_bytes)
    n = ops.num_params(cfg)
    slot_floats = 2 * ((n + 2 + 31) // 32 * 32)
    assert int(lib.inf_comm_region_bytes(C.byref(ops._desc(cfg, 0, 0)))) == 2 * 9 * slot_floats * 4 + 64 * 4
    # the vertex projection a frozen model keeps: 512 bytes per vertex, whole 128-vertex tiles
    assert ops.projection_bytes(50176) == 50176 * 512
    assert ops.projection_bytes(129) == 2 * 128 * 512
    assert ops.N_LOSS_SUMS == 3


def test_bad_descriptors_are_rejected():
    with pytest.raises(NotImplementedError):
        ops.num_params(ops.FieldConfig(k=64, hidden=256)) if False else _raise_for(ops.FieldConfig(k=64, hidden=256))
    with pytest.raises(RuntimeError):
        _raise_for(ops.FieldConfig(k=64, num_layers=6, skip=5))
    with pytest.raises(NotImplementedError):
        _raise_for(ops.FieldConfig(k=64, num_layers=12, skip=3))


def test_param_layout_reports_the_rejection_reason():
    with pytest.raises(NotImplementedError, match="embed_size=31"):
        ops.param_layout(ops.FieldConfig(k=1, num_layers=7, skip=4, view="intrinsic", view_embed=31, view_include=True,
                                         view_bneck=100, view_hidden=48))
    with pytest.raises(RuntimeError, match="skip_layer_idx"):
        ops.param_layout(ops.FieldConfig(k=64, num_layers=6, skip=5))


def _raise_for(cfg):
    lib = _lib.load()
    d = ops._desc(cfg, 0, 0)
    off, r, c = C.c_int64(), C.c_int32(), C.c_int32()
    _lib.check(lib.inf_param_info(C.byref(d), 0, C.byref(off), C.byref(r), C.byref(c)), "inf_param_info")


def test_model_state_dict_keys_match_reference():
    c = load_golden("tf_k1023")
    m = inf_b200.make_model(dict(k=list(range(c["k"])), num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3,
                                 batchnorm=False))
    sd = m.state_dict()
    assert list(sd.keys()) == c["param_names"]
    for n in sd:
        assert tuple(sd[n].shape) == tuple(c["params"][n].shape)
    m.load_state_dict(c["params"])       # reference-trained weights load unchanged
    import copy
    copy.deepcopy(m)                      # trainer.py:67


@pytest.mark.parametrize("name", ["view_intrinsic", "view_extrinsic"])
def test_view_model_state_dict_keys_match_reference(name):
    import copy
    import types
    c = load_golden(name)
    m = inf_b200.make_model(dict(k=c["k"], num_layers=c["num_layers"], mlp_hidden_dim=128, skip_layer_idx=c["skip"],
                                 view_dependence=c["view"]),
                            mesh=types.SimpleNamespace(face_normals=c["face_normals"].numpy()))
    sd = m.state_dict()
    assert list(sd.keys()) == c["param_names"]          # face_normals / freq_bands are non-persistent (model.py:146)
    assert [n for n, _ in m.named_parameters()] == c["param_names"]
    for n in sd:
        assert tuple(sd[n].shape) == tuple(c["params"][n].shape)
    m.load_state_dict(c["params"])
    copy.deepcopy(m)


def test_unsupported_options_raise():
    base = dict(k=64, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3)
    for bad in (dict(batchnorm=True), dict(activation="sine"), dict(feature_strategy="rff"), dict(mlp_hidden_dim=64)):
        with pytest.raises(NotImplementedError):
            inf_b200.make_model({**base, **bad})


def test_cpu_tensors_fail_loudly():
    import torch
    m = inf_b200.make_model(dict(k=64, num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3))
    with pytest.raises(RuntimeError):
        m({"eigenfunctions": torch.zeros(4, 64)})


def test_load_trained_model_reads_both_checkpoint_forms(tmp_path):
    # utils.py:22-29: a bare state_dict (train.py's model.pt) or the trainer's {"model_state_dict": ...} wrapper
    c = load_golden("tf_k64")
    cfg = dict(k=c["k"], num_layers=6, mlp_hidden_dim=128, skip_layer_idx=3)
    torch.save(c["params"], tmp_path / "bare.pt")
    torch.save({"model_state_dict": c["params"], "epoch": 3}, tmp_path / "wrapped.pt")
    for name in ("bare.pt", "wrapped.pt"):
        m = inf_b200.load_trained_model(cfg, str(tmp_path / name), "cpu")
        for n, t in m.state_dict().items():
            assert torch.equal(t, c["params"][n]), n
