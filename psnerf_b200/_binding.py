"""ctypes binding of libpsnerf_b200.so (the C ABI declared in include/psnerf_b200.h).

Tensors cross the boundary as raw device pointers (``tensor.data_ptr()``) plus sizes; the CUDA stream is
``torch.cuda.current_stream().cuda_stream``.  Errors come back as negative ints and are raised as RuntimeError
with psn_last_error().  There is deliberately no fallback: a missing library or a non-sm_100 device raises.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSNERF_B200_LIB") or os.path.join(_HERE, "lib", "libpsnerf_b200.so")  # env override: bring-up builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "psnerf_b200.h")

PREC_FP32, PREC_TC, PREC_TC_MIXED, PREC_TC_TWOLEVEL = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "tc": PREC_TC, "tc_mixed": PREC_TC_MIXED, "tc_two_level": PREC_TC_TWOLEVEL}
NET_GEO, NET_APP, NET_S2 = 0, 1, 2
OUT_ALPHA, OUT_NEG_LOGIT, OUT_LOGIT = 0, 1, 2

_lib = None


class MlpDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_layers", C.c_int), ("octaves", C.c_int), ("skip", C.c_int),
                ("final_act", C.c_int), ("rescale", C.c_float)]


class UnisurfParams(C.Structure):
    _fields_ = [("near_", C.c_float), ("radius", C.c_float), ("delta", C.c_float), ("tau", C.c_float),
                ("march_steps", C.c_int), ("secant_steps", C.c_int), ("steps_in", C.c_int), ("steps_out", C.c_int),
                ("white_background", C.c_int)]


class TrainNet(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("skip", C.c_int), ("final_act", C.c_int), ("in_dims", C.POINTER(C.c_int)),
                ("out_dims", C.POINTER(C.c_int)), ("W", C.POINTER(C.c_void_p)), ("b", C.POINTER(C.c_void_p)),
                ("dW", C.POINTER(C.c_void_p)), ("db", C.POINTER(C.c_void_p))]


class AdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64)]


class AdamHyper(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double),
                ("step", C.c_int64)]


class ShadeParams(C.Structure):
    _fields_ = [("n_freqs_xyz", C.c_int), ("n_freqs_normal", C.c_int), ("nbasis", C.c_int), ("specular_rgb", C.c_int),
                ("intensity_kind", C.c_int), ("intensity", C.c_float), ("render_model", C.c_int), ("fresnel_f0", C.c_float)]


def declared_symbols():
    """Every psn_* function declared in include/psnerf_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psn_[a-z0-9_]+)\s*\(", src)))


def load():
    """dlopen the library and verify that it exports every declared entry point (works without a GPU)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("psnerf_b200: %s is missing - run `python __graft_entry__.py` (build()) first; "
                           "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError("psnerf_b200: library does not export %s" % missing)
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
    lib.psn_version.restype = C.c_int
    lib.psn_last_error.restype = C.c_char_p
    lib.psn_launch_count.restype = i64
    lib.psn_profile_enable.argtypes = [i32]
    lib.psn_profile_collect.argtypes = [i32, C.POINTER(i64), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.psn_device_check.argtypes = [C.POINTER(C.c_int)]
    lib.psn_mlp_create.argtypes = [C.POINTER(MlpDesc), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(vp),
                                   C.POINTER(vp), vp, C.POINTER(vp)]
    lib.psn_mlp_free.argtypes = [vp]
    lib.psn_workspace_bytes.argtypes = [C.c_char_p, i64, i64, i64]
    lib.psn_workspace_bytes.restype = i64
    lib.psn_occupancy.argtypes = [vp, vp, i64, i32, vp, i32, vp]
    lib.psn_infer_occ.argtypes = [vp, vp, i64, vp, i32, vp]
    lib.psn_gradient.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp]
    lib.psn_radiance.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, i64, i32, vp]
    lib.psn_rays_from_pixels.argtypes = [vp, i64, C.POINTER(f32), i32, vp, vp]
    lib.psn_raymarch.argtypes = [vp, C.POINTER(f32), vp, i64, f32, f32, i32, i32, f32, vp, vp, i64, i32, vp]
    lib.psn_render_unisurf.argtypes = [vp, vp, C.POINTER(f32), vp, i64, C.POINTER(UnisurfParams), vp, vp, vp, vp, vp,
                                       vp, vp, vp, i64, i32, vp]
    lib.psn_shadow_visibility.argtypes = [vp, vp, vp, i64, i32, f32, f32, i32, f32, vp, vp, i64, i32, vp]
    lib.psn_shade_stage2.argtypes = [vp, vp, vp, vp, vp, C.POINTER(ShadeParams), vp, vp, vp, vp, i64, i64, vp, i32,
                                     vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp]
    lib.psn_shade_stage2_edit.argtypes = [vp, vp, vp, vp, vp, C.POINTER(ShadeParams), vp, vp, vp, vp, i64, i64, vp, i32,
                                          vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp]
    lib.psn_s2_point_nets.argtypes = [vp, vp, i32, vp, i64, vp, vp, i32, i32, vp]
    lib.psn_s2_visibility.argtypes = [vp, i32, vp, i64, vp, i32, vp, vp, i64, i32, vp]
    lib.psn_tc_debug_layer.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.psn_tc_debug_trace.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.psn_tc_debug_trace_q.argtypes = [vp, vp, i64, vp, vp, i32, vp]
    lib.psn_tc_debug_trace_rad.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, i32, vp]
    lib.psn_tc_gemm_debug.argtypes = [i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, i64, i32, vp]
    lib.psn_tc_gemm_debug_fused.argtypes = [i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, i64, i32, vp, vp, vp, i64, f32, vp]
    tn = C.POINTER(TrainNet)
    lib.psn_s1_train_tape_bytes.argtypes = [tn, tn, i32, i32, i64]
    lib.psn_s1_train_tape_bytes.restype = i64
    lib.psn_s1_train_ws_bytes.argtypes = [tn, tn, i32, i32, i64]
    lib.psn_s1_train_ws_bytes.restype = i64
    lib.psn_s1_train_forward.argtypes = [tn, tn, i32, i32, f32, vp, vp, i64, vp, vp, vp, vp, i64, vp, i64, vp]
    lib.psn_s1_train_backward.argtypes = [tn, tn, i32, i32, f32, i64, vp, vp, vp, vp, i64, vp, i64, vp]
    lib.psn_composite_bwd.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp, vp]
    lib.psn_s2_train_tape_bytes.argtypes = [tn, tn, tn, tn, i64, i32, i32]
    lib.psn_s2_train_tape_bytes.restype = i64
    lib.psn_s2_train_forward.argtypes = [tn, tn, tn, tn, vp, vp, C.POINTER(ShadeParams), vp, vp, vp, i64, i64, vp, i32, vp, vp, vp, i32,
                                         vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, i64, i32, vp]
    lib.psn_s2_train_backward.argtypes = [tn, tn, tn, tn, vp, C.POINTER(ShadeParams), vp, vp, i64, i64, vp, i32, vp, i32,
                                          vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp]
    lib.psn_composite.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
    lib.psn_adam_step.argtypes = [C.POINTER(AdamTensor), i32, C.POINTER(AdamHyper), vp]
    lib.psn_sparse_adam_step.argtypes = [vp, vp, vp, i64, i32, vp, vp, i64, C.POINTER(AdamHyper), vp]
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().psn_last_error()
        raise RuntimeError("psnerf_b200 %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def require_device():
    """Raise unless a CUDA sm_100 device is usable (no CPU fallback by design)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("psnerf_b200: no CUDA device - the hot path has no CPU fallback")
    n = C.c_int(0)
    check(load().psn_device_check(C.byref(n)), "psn_device_check")
    return n.value
