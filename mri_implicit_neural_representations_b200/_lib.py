"""ctypes binding of libinr_b200.so (C ABI declared in include/inr_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at import time."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libinr_b200.so")


class InrError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("in_features", C.c_int32), ("out_features", C.c_int32),
                ("depth", C.c_int32), ("width", C.c_int32), ("last_act", C.c_int32),
                ("encoder", C.c_int32), ("enc_size", C.c_int32), ("w0", C.c_float),
                ("hidden_omega_0", C.c_float), ("sigma0", C.c_float), ("head_mask", C.c_int32), ("bounds", C.c_float * 20)]


class LossDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("hdr_eps", C.c_float), ("hdr_sigma", C.c_float), ("hdr_factor", C.c_float),
                ("tv_weight", C.c_float), ("tv_h", C.c_int32), ("tv_w", C.c_int32),
                ("dp_norm", C.c_void_p), ("dp_rows", C.c_int32), ("cons_weight", C.c_float), ("cons_bounds", C.c_float * 16)]


class TensorInfo(C.Structure):
    _fields_ = [("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32),
                ("layer", C.c_int32), ("is_bias", C.c_int32), ("is_complex", C.c_int32), ("frozen", C.c_int32)]


MODEL = {"SIREN": 1, "FFN": 2, "WIRE": 3, "Fourier": 4, "MultiscaleFourier": 5, "BoundedFourier": 6, "Gabor": 7, "KGabor": 7, "WIRE2D": 8}
ENC = {"none": 0, "gauss": 1, "LogF": 2}
LAST = {"linear": 0, "tanh": 1, "sigmoid": 2, "sin": 3}
LOSS = {"none": 0, "L2": 1, "L1": 2, "MSLE": 3, "tanh": 4, "LSL": 5, "HDR": 6}

# every symbol include/inr_b200.h declares (tests check the library exports all of them)
EXPORTS = ["inr_last_error", "inr_plan_create", "inr_plan_destroy", "inr_plan_param_count",
           "inr_plan_tensor_count", "inr_plan_tensor", "inr_wpack_bytes", "inr_workspace_bytes",
           "inr_scalars_offset", "inr_workspace_layout", "inr_pack_weights", "inr_forward", "inr_backward",
           "inr_forward_dist", "inr_backward_dist", "inr_adam_step", "inr_adam_step_peers",
           "inr_train_step", "inr_train_step_dist", "inr_grad_step", "inr_grad_step_dist", "inr_profile_step", "inr_debug_set_trace", "inr_selftest_umma", "inr_set_sm_budget"]


def _load():
    if not os.path.exists(LIB_PATH):
        raise InrError(
            f"{LIB_PATH} is missing: build it with `python -m mri_implicit_neural_representations_b200.build` "
            "(nvcc, sm_100a). This engine has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.inr_last_error.restype = C.c_char_p
    lib.inr_plan_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    lib.inr_plan_destroy.argtypes = [vp]
    lib.inr_plan_param_count.argtypes = [vp, C.POINTER(i64)]
    lib.inr_plan_tensor_count.argtypes = [vp, C.POINTER(i32)]
    lib.inr_plan_tensor.argtypes = [vp, i32, C.POINTER(TensorInfo)]
    lib.inr_wpack_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.inr_workspace_bytes.argtypes = [vp, i64, C.POINTER(C.c_size_t)]
    lib.inr_scalars_offset.argtypes = [vp, i64, C.POINTER(C.c_size_t)]
    lib.inr_workspace_layout.argtypes = [vp, i64, C.POINTER(C.c_uint64), i32]
    lib.inr_pack_weights.argtypes = [vp, vp, vp, vp]
    lib.inr_forward.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, i32, vp]
    lib.inr_backward.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    lib.inr_forward_dist.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, vp, i32, vp]
    lib.inr_backward_dist.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp]
    lib.inr_adam_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.inr_train_step.argtypes = [vp, C.POINTER(LossDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp,
                                   vp, vp, vp]
    lib.inr_train_step_dist.argtypes = [vp, C.POINTER(LossDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp,
                                        vp, vp, vp]
    lib.inr_grad_step_dist.argtypes = [vp, C.POINTER(LossDesc), vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.inr_grad_step.argtypes = [vp, C.POINTER(LossDesc), vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.inr_profile_step.argtypes = [vp, C.POINTER(LossDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i32,
                                     C.POINTER(C.c_float), vp]
    lib.inr_adam_step_peers.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), i32, i32, vp, vp, vp, vp, vp, vp]
    lib.inr_debug_set_trace.argtypes = [vp]
    lib.inr_set_sm_budget.argtypes = [i32]
    lib.inr_selftest_umma.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for name in EXPORTS:
        if name != "inr_last_error":
            getattr(lib, name).restype = C.c_int
    return lib


lib = _load()


def check(rc: int, what: str = ""):
    if rc != 0:
        raise InrError(f"{what} failed ({rc}): {lib.inr_last_error().decode()}")
