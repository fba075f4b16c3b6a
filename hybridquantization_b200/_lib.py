"""ctypes binding of libhq_b200.so (the C ABI declared in include/hq_b200.h).

There is no Python or CPU fallback: if the shared library is missing, import fails loudly
with the build command; if no GPU is present, hq_create() fails and HqError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HQ_B200_LIB selects another build of the SAME library (kernel experiments); there is still no fallback
LIB_PATH = os.environ.get("HQ_B200_LIB") or os.path.join(_HERE, "libhq_b200.so")

HQ_OK = 0
ERR_NAMES = {1: "HQ_ERR_INVALID", 2: "HQ_ERR_CUDA", 3: "HQ_ERR_NO_IMAGE", 4: "HQ_ERR_UNSUPPORTED", 5: "HQ_ERR_CALLBACK"}
WHITEPOINT_D65, WHITEPOINT_D50 = 0, 1
SPACE_LAB, SPACE_SRGB = 0, 1
COST_LAB, COST_SCIELAB = 0, 1
EVAL_SUMS, EVAL_FORCE_DIRECT, EVAL_FORCE_CHUNKED, EVAL_FORCE_PREFILTER, EVAL_PRUNE, EVAL_ALLREDUCE = 1, 2, 4, 8, 16, 32
PRUNE_OFF, PRUNE_AUTO, PRUNE_ON = 0, 1, 2
MAX_COLORS = 1024
MAX_COLORS_PRUNED = 4096
MAX_COLORS_ANY = 1 << 24
COMM_ID_BYTES = 128
PEER_HANDLE_BYTES = 64
DELTAE_CIE76, DELTAE_CIE94, DELTAE_CIEDE2000 = 0, 1, 2
ERR_FX_NAN = -(1 << 63)


class HqError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class SwasaParams(C.Structure):
    """hq_swasa_params (HybridQuantization.java:197-224 + space, seed)."""

    _fields_ = [
        ("population", C.c_int), ("imax", C.c_int), ("iTc", C.c_int), ("delta", C.c_float),
        ("convergence", C.c_int), ("conv_delay", C.c_float), ("conv_spread", C.c_float),
        ("t0", C.c_float), ("alpha", C.c_float), ("s0", C.c_float), ("beta", C.c_float),
        ("space", C.c_int), ("seed", C.c_int64), ("cost_model", C.c_int),
    ]


class JavaRandomState(C.Structure):
    _fields_ = [("state", C.c_uint64)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_double)

# every symbol include/hq_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "hq_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "hq_destroy": (None, [_P]),
    "hq_last_error": (C.c_char_p, [_P]),
    "hq_device_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "hq_set_delta_e": (C.c_int, [_P, C.c_int]),
    "hq_set_image_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "hq_set_image_u8_sharded": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hq_set_image_f32_planar": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "hq_set_image_f32_planar_sharded": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 7),
    "hq_set_image_u8_device": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "hq_get_lab": (C.c_int, [_P, _P]),
    "hq_image_pixels": (C.c_uint64, [_P]),
    "hq_eval_palettes": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "hq_result_words": (C.c_int, [C.c_int, C.c_int]),
    "hq_eval_palettes_device": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hq_cost": (C.c_double, [C.c_int64, _P, C.c_int, C.c_uint64, C.c_float]),
    "hq_quantize": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "hq_scielab_configure": (C.c_int, [_P, C.c_int, C.c_float]),
    "hq_scielab_set_filters": (C.c_int, [_P, _P, _P, C.c_int]),
    "hq_scielab_get_filters": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int)]),
    "hq_scielab_get_image": (C.c_int, [_P, _P]),
    "hq_error_image": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_double)]),
    "hq_error_image_f32_planar": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "hq_rgb_to_xyz": (C.c_int, [_P, _P, _P, _P, C.c_size_t, _P]),
    "hq_xyz_to_scielab": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "hq_scielab_set_image": (C.c_int, [_P, _P]),
    "hq_delta_e_images": (C.c_int, [_P, _P, _P, C.c_size_t, _P, C.POINTER(C.c_double)]),
    "hq_scielab_force_generic": (C.c_int, [_P, C.c_int]),
    "hq_scielab_build_filters": (C.c_int, [C.c_int, C.c_float, _P, _P, C.POINTER(C.c_int)]),
    "hq_eval_palettes_scielab": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hq_set_allreduce": (C.c_int, [_P, ALLREDUCE_FN, _P]),
    "hq_comm_get_unique_id": (C.c_int, [_P]),
    "hq_comm_init_rank": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "hq_comm_allreduce": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "hq_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hq_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "hq_multi_device_count": (C.c_int, [_P]),
    "hq_swasa_default_params": (None, [C.POINTER(SwasaParams)]),
    "hq_find_best_quantization": (C.c_int, [_P, C.c_int, C.POINTER(SwasaParams), C.c_uint64, _P, C.POINTER(C.c_double), _P, C.POINTER(C.c_int)]),
    "hq_request_stop": (None, [_P]),
    "hq_set_progress": (C.c_int, [_P, PROGRESS_FN, _P]),
    "hq_set_pruning": (C.c_int, [_P, C.c_int]),
    "hq_set_graphs": (C.c_int, [_P, C.c_int]),
    "hq_search_eval_flags": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "hq_pruning_stats": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_double)]),
    "hq_java_random_seed": (None, [C.POINTER(JavaRandomState), C.c_int64]),
    "hq_java_random_next": (C.c_int32, [C.POINTER(JavaRandomState), C.c_int]),
    "hq_java_random_next_float": (C.c_float, [C.POINTER(JavaRandomState)]),
    "hq_java_random_next_double": (C.c_double, [C.POINTER(JavaRandomState)]),
    "hq_swasa_generate_random_colors": (None, [C.POINTER(JavaRandomState), C.c_int, _P]),
    "hq_swasa_generate_neighboring_colors": (None, [C.POINTER(SwasaParams), C.POINTER(JavaRandomState), _P, _P, C.c_int, C.c_int]),
    "hq_swasa_max_step_width": (C.c_float, [C.POINTER(SwasaParams), C.c_int]),
    "hq_comm_peer_handle": (C.c_int, [_P, _P]),
    "hq_comm_open_peers": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "hq_comm_close_peers": (None, [_P]),
    "hq_comm_peers_open": (C.c_int, [_P]),
    "hq_set_profiling": (C.c_int, [_P, C.c_int]),
    "hq_last_assign_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "hq_last_rgb_to_lab_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "hq_last_scielab_stage_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "hq_measure_fp32_peak": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "hq_host_math_range": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, _P, C.c_int]),
    "hq_device_math_range": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32, _P]),
    "hq_host_srgb_to_lab": (None, [_P, C.c_int, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the CUDA extension; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is required (there is no CPU fallback). "
            "Build it with `python -m hybridquantization_b200.build` or `python -c 'import __graft_entry__ as g; g.build()'`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc: int) -> None:
    if rc != HQ_OK:
        msg = load().hq_last_error(ctx)
        raise HqError(rc, msg.decode() if msg else "")
