"""ctypes binding of libinf_field.so (C ABI declared in include/inf_field.h).

There is no CPU fallback: if the library has not been built the import of any compute entry point
raises, and every compute call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libinf_field.so")

ABI_VERSION = 6
IDX_I32, IDX_I64 = 0, 1
TABLE_F32, TABLE_F16 = 0, 1
INPUT_GATHER, INPUT_DENSE = 0, 1
VIEW_NONE, VIEW_INTRINSIC, VIEW_EXTRINSIC = 0, 1, 2
PATH_AUTO, PATH_RAY, PATH_VERTEX = 0, 1, 2
PROFILE_STAGES = 6
MAX_PARAMS = 20
ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = -1, -2, -3, -4


class InfDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("idx_dtype", C.c_int32),
        ("n_rays", C.c_int64), ("n_rows", C.c_int64),
        ("k", C.c_int32), ("k_pad", C.c_int32), ("hidden", C.c_int32), ("num_layers", C.c_int32),
        ("skip", C.c_int32), ("table_dtype", C.c_int32), ("input_mode", C.c_int32),
        ("view_strategy", C.c_int32), ("view_embed", C.c_int32), ("view_include", C.c_int32),
        ("view_bneck", C.c_int32), ("view_hidden", C.c_int32),
        ("n_faces", C.c_int64), ("n_src_rays", C.c_int64), ("ray_idx_dtype", C.c_int32), ("path", C.c_int32),
        ("grad_scale", C.c_float), ("reserved0", C.c_float),
    ]


class InfTensors(C.Structure):
    _fields_ = [
        ("table", C.c_void_p), ("vids", C.c_void_p), ("bary", C.c_void_p), ("target", C.c_void_p),
        ("grad_pred", C.c_void_p), ("unit_ray_dirs", C.c_void_p), ("face_idxs", C.c_void_p),
        ("face_normals", C.c_void_p), ("ray_idxs", C.c_void_p), ("table_image", C.c_void_p), ("col_exp", C.c_void_p),
        ("packed", C.c_void_p),
        ("pred", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("proj", C.c_void_p),
    ]


class InfGrads(C.Structure):
    _fields_ = [("flat", C.c_void_p), ("loss_sums", C.c_void_p)]


class InfAdam(C.Structure):
    _fields_ = [("params", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("packed", C.c_void_p),
                ("col_exp", C.c_void_p), ("step_dev", C.c_void_p), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("step", C.c_int32), ("flags", C.c_int32)]


ADAM_STEP_AUTOINC = 1


MAX_PEERS = 8
IPC_HANDLE_BYTES = 64


class InfComm(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("regions", C.c_void_p * MAX_PEERS), ("region_bytes", C.c_size_t)]


EXPORTS = {
    # name: (restype, argtypes)
    "inf_version": (C.c_int, []),
    "inf_last_error": (C.c_char_p, []),
    "inf_launch_count": (C.c_longlong, []),
    "inf_debug_flags": (C.c_int, []),
    "inf_num_params": (C.c_int64, [C.POINTER(InfDesc)]),
    "inf_num_param_tensors": (C.c_int32, [C.POINTER(InfDesc)]),
    "inf_param_info": (C.c_int, [C.POINTER(InfDesc), C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32)]),
    "inf_packed_bytes": (C.c_size_t, [C.POINTER(InfDesc)]),
    "inf_workspace_bytes": (C.c_size_t, [C.POINTER(InfDesc), C.c_int]),
    "inf_relayout_table": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "inf_column_exponents": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "inf_relayout_table_scaled": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_int32, C.c_void_p]),
    "inf_pack_weights_scaled": (C.c_int, [C.POINTER(InfDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p]),
    "inf_gather_interpolate_scaled": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "inf_table_image_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "inf_build_table_image": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "inf_project_vertices": (C.c_int, [C.POINTER(InfDesc), C.POINTER(InfTensors), C.c_void_p, C.c_void_p]),
    "inf_projection_bytes": (C.c_size_t, [C.c_int64]),
    "inf_project_vertices_raw": (C.c_int, [C.POINTER(InfDesc), C.POINTER(InfTensors), C.c_void_p, C.c_void_p]),
    "inf_uses_vertex_path": (C.c_int, [C.POINTER(InfDesc), C.POINTER(InfTensors)]),
    "inf_pack_weights": (C.c_int, [C.POINTER(InfDesc), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "inf_forward": (C.c_int, [C.POINTER(InfDesc), C.POINTER(InfTensors), C.c_void_p]),
    "inf_forward_backward": (C.c_int, [C.POINTER(InfDesc), C.POINTER(InfTensors), C.POINTER(InfGrads), C.c_void_p]),
    "inf_train_step": (C.c_int, [C.POINTER(InfDesc), C.POINTER(InfTensors), C.POINTER(InfGrads), C.POINTER(InfAdam), C.c_void_p]),
    "inf_train_step_comm": (C.c_int, [C.POINTER(InfDesc), C.POINTER(InfTensors), C.POINTER(InfGrads), C.POINTER(InfAdam),
                                      C.POINTER(InfComm), C.c_void_p]),
    "inf_comm_region_bytes": (C.c_size_t, [C.POINTER(InfDesc)]),
    "inf_comm_create": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "inf_comm_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "inf_comm_close": (C.c_int, [C.c_void_p]),
    "inf_comm_free": (C.c_int, [C.c_void_p]),
    "inf_adam_pack_step": (C.c_int, [C.POINTER(InfDesc), C.c_void_p, C.POINTER(InfAdam), C.c_void_p]),
    "inf_gather_interpolate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p]),
    "inf_profile_enable": (C.c_int, [C.c_int]),
    "inf_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                   C.POINTER(C.c_int)]),
    "inf_profile_read_stages": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "inf_debug_umma": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library once; raise loudly if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C intrinsic-neural-fields_b200/csrc`). The B200 path has no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.inf_version() != ABI_VERSION:
        raise RuntimeError(f"libinf_field.so ABI {lib.inf_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = load().inf_last_error().decode("utf-8", "replace")
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed ({rc}): {msg}")
