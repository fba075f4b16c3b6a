"""ctypes binding of the CPU oracle (oracle/libhq_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under hybridquantization_b200/ imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhq_oracle.so")
WHITE_D65, WHITE_D50 = 0, 1
SPACE_LAB, SPACE_SRGB = 0, 1


class SwasaParams(C.Structure):
    _fields_ = [("population", C.c_int), ("imax", C.c_int), ("iTc", C.c_int), ("delta", C.c_float),
                ("convergence", C.c_int), ("conv_delay", C.c_float), ("conv_spread", C.c_float),
                ("t0", C.c_float), ("alpha", C.c_float), ("s0", C.c_float), ("beta", C.c_float),
                ("whitepoint", C.c_int), ("space", C.c_int), ("seed", C.c_int64),
                ("cost_model", C.c_int), ("dpi", C.c_int), ("viewing_distance", C.c_float)]


class Rng(C.Structure):
    _fields_ = [("state", C.c_uint64)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("hq_oracle.c", "hq_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "libhq_oracle.so"], check=True)
    return LIB_PATH


_lib = None
_P = C.c_void_p


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.hqo_u8_to_unit.restype = C.c_float; L.hqo_u8_to_unit.argtypes = [C.c_uint]
        for n in ("hqo_srgb_decode", "hqo_pow_2p4", "hqo_cbrt_pow"):
            getattr(L, n).restype = C.c_float; getattr(L, n).argtypes = [C.c_float]
        L.hqo_lab_constants.restype = C.c_float; L.hqo_lab_constants.argtypes = [C.c_int]
        L.hqo_srgb_to_lab.argtypes = [_P, C.c_int, _P]
        L.hqo_srgb_to_opp.argtypes = [_P, _P]
        L.hqo_opp_to_lab.argtypes = [_P, C.c_int, _P]
        L.hqo_image_planes.argtypes = [_P, C.c_size_t, C.c_int] + [_P] * 6 + [C.c_int]
        L.hqo_assign_reduce.argtypes = [_P, C.c_size_t, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int]
        L.hqo_assign_reduce_planes.argtypes = [_P, _P, C.c_size_t, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int]
        L.hqo_cost.restype = C.c_double; L.hqo_cost.argtypes = [C.c_int64, _P, C.c_int, C.c_uint64, C.c_float]
        L.hqo_quantize.argtypes = [_P, C.c_size_t, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int]
        L.hqo_rng_seed.argtypes = [C.POINTER(Rng), C.c_int64]
        L.hqo_rng_next.restype = C.c_int32; L.hqo_rng_next.argtypes = [C.POINTER(Rng), C.c_int]
        L.hqo_rng_next_float.restype = C.c_float; L.hqo_rng_next_float.argtypes = [C.POINTER(Rng)]
        L.hqo_rng_next_double.restype = C.c_double; L.hqo_rng_next_double.argtypes = [C.POINTER(Rng)]
        L.hqo_swasa_defaults.argtypes = [C.POINTER(SwasaParams)]
        L.hqo_generate_random_colors.argtypes = [C.POINTER(Rng), C.c_int, _P]
        L.hqo_max_step_width.restype = C.c_float; L.hqo_max_step_width.argtypes = [C.POINTER(SwasaParams), C.c_int]
        L.hqo_generate_neighboring_colors.argtypes = [C.POINTER(SwasaParams), C.POINTER(Rng), _P, _P, C.c_int, C.c_int]
        L.hqo_find_best_quantization.restype = C.c_double
        L.hqo_find_best_quantization.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.POINTER(SwasaParams), _P, _P, C.c_int]
        L.hqo_synth_image.argtypes = [_P, C.c_int, C.c_int, C.c_uint64, C.c_int]
        L.hqo_math_range.argtypes = [C.c_int, C.c_uint32, C.c_uint32, _P, C.c_int]
        L.hqo_scielab_filters.restype = C.c_int; L.hqo_scielab_filters.argtypes = [C.c_int, C.c_double, _P, _P, C.c_int]
        L.hqo_scielab_image.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int]
        L.hqo_scielab_eval.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int]
        L.hqo_image_planes_f32.argtypes = [_P, C.c_size_t, C.c_int, _P, C.c_int]
        L.hqo_scielab_image_f32.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int]
        L.hqo_scielab_eval_planes.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int]
        L.hqo_delta_e94_array.argtypes = [_P, _P, C.c_size_t, _P]
        L.hqo_error_image_f32.restype = C.c_double
        L.hqo_error_image_f32.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, C.c_int]
        L.hqo_error_image.restype = C.c_double
        L.hqo_error_image.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def default_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def srgb_to_lab(rgb, whitepoint=WHITE_D65) -> np.ndarray:
    a = np.ascontiguousarray(rgb, np.float32); out = np.empty(3, np.float32)
    load().hqo_srgb_to_lab(_ptr(a), whitepoint, _ptr(out))
    return out


def image_planes(rgb_u8: np.ndarray, whitepoint=WHITE_D65, threads=None):
    """returns (unit [3,n], lab [3,n]) float32"""
    rgb = np.ascontiguousarray(rgb_u8, np.uint8).reshape(-1, 3)
    n = rgb.shape[0]
    unit = np.empty((3, n), np.float32); lab = np.empty((3, n), np.float32)
    load().hqo_image_planes(_ptr(rgb), n, whitepoint, _ptr(unit[0]), _ptr(unit[1]), _ptr(unit[2]),
                            _ptr(lab[0]), _ptr(lab[1]), _ptr(lab[2]), threads or default_threads())
    return unit, lab


def image_planes_f32(planes: np.ndarray, whitepoint=WHITE_D65, threads=None):
    """planes: float32 sRGB in [0,1], [3, rows, width] or [3, n] (im.getDataXYCAsFloat()); returns (unit [3,n], lab [3,n])"""
    unit = np.ascontiguousarray(planes, np.float32).reshape(3, -1)
    lab = np.empty_like(unit)
    load().hqo_image_planes_f32(_ptr(unit), unit.shape[1], whitepoint, _ptr(lab), threads or default_threads())
    return unit, lab


def assign_reduce_planes(unit, lab, palettes, space=SPACE_LAB, whitepoint=WHITE_D65, want_idx=False, threads=None):
    palettes = np.ascontiguousarray(palettes, np.float32)
    if palettes.ndim == 2:
        palettes = palettes[None]
    B, K, _ = palettes.shape
    n = lab.shape[1]
    err = np.empty(B, np.int64); counts = np.empty((B, K), np.uint64); sums = np.empty((B, K, 3), np.int64)
    idx = np.empty((B, n), np.uint16) if want_idx else None
    unit = None if unit is None else np.ascontiguousarray(unit, np.float32)
    lab = np.ascontiguousarray(lab, np.float32)
    load().hqo_assign_reduce_planes(_ptr(unit), _ptr(lab), n, whitepoint, _ptr(palettes), B, K, space, _ptr(err), _ptr(counts),
                                    _ptr(sums), _ptr(idx), threads or default_threads())
    return {"err_fx": err, "counts": counts, "sums_fx": sums, "idx": idx}


def assign_reduce(rgb_u8, palettes, space=SPACE_LAB, whitepoint=WHITE_D65, want_idx=False, threads=None):
    unit, lab = image_planes(rgb_u8, whitepoint, threads)
    return assign_reduce_planes(unit, lab, palettes, space, whitepoint, want_idx, threads)


def cost(err_fx, counts, n_total, delta) -> float:
    c = np.ascontiguousarray(counts, np.uint64)
    return load().hqo_cost(int(err_fx), _ptr(c), c.shape[0], n_total, delta)


def quantize(rgb_u8, palette, space=SPACE_LAB, whitepoint=WHITE_D65, threads=None):
    rgb = np.ascontiguousarray(rgb_u8, np.uint8).reshape(-1, 3)
    n = rgb.shape[0]
    palette = np.ascontiguousarray(palette, np.float32)
    out = np.empty((n, 3), np.uint8); f32 = np.empty((n, 4), np.float32); idx = np.empty(n, np.uint16)
    load().hqo_quantize(_ptr(rgb), n, whitepoint, _ptr(palette), palette.shape[0], space, _ptr(out), _ptr(f32), _ptr(idx),
                        threads or default_threads())
    return {"rgb": out, "f32": f32, "idx": idx}


def swasa_params(**kw) -> SwasaParams:
    p = SwasaParams(); load().hqo_swasa_defaults(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def find_best_quantization(rgb_u8, K, params: SwasaParams, trace=False, threads=None):
    rgb = np.ascontiguousarray(rgb_u8, np.uint8)
    h, w = rgb.shape[:2]
    best = np.empty((K, 4), np.float32)
    tr = np.empty((params.imax + 1, params.population), np.float64) if trace else None
    err = load().hqo_find_best_quantization(_ptr(rgb), w, h, K, C.byref(params), _ptr(best), _ptr(tr), threads or default_threads())
    return best, err, tr


def synth_image(width, height, seed, smooth=False) -> np.ndarray:
    out = np.empty((height, width, 3), np.uint8)
    load().hqo_synth_image(_ptr(out), width, height, seed, int(smooth))
    return out


def math_range(which: int, first_bits: int, count: int, threads=None) -> np.ndarray:
    out = np.empty(count, np.float32)
    load().hqo_math_range(which, first_bits, count, _ptr(out), threads or default_threads())
    return out


def scielab_filters(dpi=72, viewing_distance=45.0, max_taps=4096):
    """(filters [7, taps], abs3 [taps]) of ScielabProcessor.java:66-181"""
    f = np.zeros(7 * max_taps, np.float32); a = np.zeros(max_taps, np.float32)
    t = load().hqo_scielab_filters(dpi, float(viewing_distance), _ptr(f), _ptr(a), max_taps)
    if t < 0:
        raise ValueError("too many taps")
    return f[:7 * t].reshape(7, t).copy(), a[:t].copy()


def scielab_image(rgb_u8, filters, abs3, whitepoint=WHITE_D65, threads=None) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb_u8, np.uint8)
    h, w = rgb.shape[:2]
    lab = np.empty((3, h * w), np.float32)
    filters = np.ascontiguousarray(filters, np.float32); abs3 = np.ascontiguousarray(abs3, np.float32)
    load().hqo_scielab_image(_ptr(rgb), w, h, whitepoint, _ptr(filters), _ptr(abs3), filters.shape[1], _ptr(lab), threads or default_threads())
    return lab


def scielab_eval(rgb_u8, filters, abs3, scielab_orig, palettes, space=SPACE_SRGB, whitepoint=WHITE_D65, threads=None):
    rgb = np.ascontiguousarray(rgb_u8, np.uint8)
    h, w = rgb.shape[:2]
    palettes = np.ascontiguousarray(palettes, np.float32)
    if palettes.ndim == 2:
        palettes = palettes[None]
    B, K, _ = palettes.shape
    err = np.empty(B, np.int64); counts = np.empty((B, K), np.uint64)
    filters = np.ascontiguousarray(filters, np.float32); abs3 = np.ascontiguousarray(abs3, np.float32)
    so = np.ascontiguousarray(scielab_orig, np.float32)
    load().hqo_scielab_eval(_ptr(rgb), w, h, whitepoint, _ptr(filters), _ptr(abs3), filters.shape[1], _ptr(so), _ptr(palettes), B, K, space,
                            _ptr(err), _ptr(counts), threads or default_threads())
    return {"err_fx": err, "counts": counts}


def scielab_image_f32(planes, filters, abs3, whitepoint=WHITE_D65, threads=None) -> np.ndarray:
    """planes: float32 [3, rows, width]"""
    unit = np.ascontiguousarray(planes, np.float32)
    _, h, w = unit.shape
    lab = np.empty((3, h * w), np.float32)
    filters = np.ascontiguousarray(filters, np.float32); abs3 = np.ascontiguousarray(abs3, np.float32)
    load().hqo_scielab_image_f32(_ptr(unit), w, h, whitepoint, _ptr(filters), _ptr(abs3), filters.shape[1], _ptr(lab), threads or default_threads())
    return lab


def scielab_eval_f32(planes, filters, abs3, scielab_orig, palettes, space=SPACE_SRGB, whitepoint=WHITE_D65, threads=None):
    unit = np.ascontiguousarray(planes, np.float32)
    _, h, w = unit.shape
    _, lab = image_planes_f32(unit, whitepoint, threads)
    palettes = np.ascontiguousarray(palettes, np.float32)
    if palettes.ndim == 2:
        palettes = palettes[None]
    B, K, _ = palettes.shape
    err = np.empty(B, np.int64); counts = np.empty((B, K), np.uint64)
    filters = np.ascontiguousarray(filters, np.float32); abs3 = np.ascontiguousarray(abs3, np.float32)
    so = np.ascontiguousarray(scielab_orig, np.float32)
    load().hqo_scielab_eval_planes(_ptr(unit), _ptr(lab), w, h, whitepoint, _ptr(filters), _ptr(abs3), filters.shape[1], _ptr(so), _ptr(palettes),
                                   B, K, space, _ptr(err), _ptr(counts), threads or default_threads())
    return {"err_fx": err, "counts": counts}


def error_image(rgb_a, rgb_b, filters, abs3, whitepoint=WHITE_D65, threads=None):
    a = np.ascontiguousarray(rgb_a, np.uint8); b = np.ascontiguousarray(rgb_b, np.uint8)
    h, w = a.shape[:2]
    emap = np.empty(h * w, np.float32); e8 = np.empty(h * w, np.uint8)
    filters = np.ascontiguousarray(filters, np.float32); abs3 = np.ascontiguousarray(abs3, np.float32)
    mean = load().hqo_error_image(_ptr(a), _ptr(b), w, h, whitepoint, _ptr(filters), _ptr(abs3), filters.shape[1], _ptr(emap), _ptr(e8),
                                  threads or default_threads())
    return {"deltaE": mean, "errorImage": emap, "errorImageU8": e8}


def error_image_f32(planes_a, planes_b, filters, abs3, whitepoint=WHITE_D65, threads=None):
    """both images as float planes [3, rows, width]"""
    a = np.ascontiguousarray(planes_a, np.float32); b = np.ascontiguousarray(planes_b, np.float32)
    _, h, w = a.shape
    emap = np.empty(h * w, np.float32); e8 = np.empty(h * w, np.uint8)
    filters = np.ascontiguousarray(filters, np.float32); abs3 = np.ascontiguousarray(abs3, np.float32)
    mean = load().hqo_error_image_f32(_ptr(a), _ptr(b), w, h, whitepoint, _ptr(filters), _ptr(abs3), filters.shape[1], _ptr(emap), _ptr(e8),
                                      threads or default_threads())
    return {"deltaE": mean, "errorImage": emap, "errorImageU8": e8}


def delta_e94(lab1, lab2) -> np.ndarray:
    """CIE94 branch of the reference's CIEDE kernel (cl:217-226) on [n, 3] Lab arrays; NaN where the reference's is"""
    a = np.ascontiguousarray(lab1, np.float32).reshape(-1, 3); b = np.ascontiguousarray(lab2, np.float32).reshape(-1, 3)
    out = np.empty(a.shape[0], np.float32)
    load().hqo_delta_e94_array(_ptr(a), _ptr(b), a.shape[0], _ptr(out))
    return out
