"""B200-native (sm_100a) implementation of the intrinsic-neural-field evaluation path:
gather 3 eigenfunction rows per ray -> barycentric interpolation -> MLP -> RGB, forward and
training backward, behind the call surface of tum-vision/intrinsic-neural-fields
(model.py / ray_dataloader.py / renderer.py / trainer.py; bake.py and ray_store.py for the baking call site and
the preprocessed-ray files).  Import as ``inf_b200``."""
from . import _lib
from .ops import EigenTable, FieldConfig, pack_weights, param_layout, num_params, debug_umma, kernel_error_word
from .model import TextureField, TextureFieldWithViewDependency, make_model, init_weights, load_trained_model

__all__ = ["EigenTable", "FieldConfig", "TextureField", "TextureFieldWithViewDependency", "make_model", "init_weights", "load_trained_model", "pack_weights",
           "param_layout", "num_params", "debug_umma", "kernel_error_word"]
