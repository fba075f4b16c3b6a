"""oracle/_ref — the reference's OWN sources compiled for the CPU (oracle/ref_build/build_ref.sh) and the
host sequencing that ImageManipulation.java performs between its OpenCL kernels.  TEST INFRASTRUCTURE ONLY:
imported by tests/, tests/golden/make_ref_golden.py and bench.py's cpu_baseline / --impl reference legs.
Nothing under hybridquantization_b200/ imports this.

libhq_ref.so holds, compiled from where they lie under /root/reference: every kernel of
OptimizedConvolution.cl, the SWASA class, the Java CPU colour helpers (ScielabProcessor.java:279-311) and the
annealing loop of findBestQuantization (ImageManipulation.java:490-545).  What is restated HERE, because it is
JavaCL buffer plumbing that cannot be compiled without the un-vendored jars, is only the ORDER in which the
reference enqueues those kernels (cited per function) and the host mean of ImageManipulation.java:736-768.

Citations: File:line under /root/reference/src/plugins/dbrasseur/hybridquantization/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libhq_ref.so")
BUILD_SCRIPT = os.path.join(_HERE, "ref_build", "build_ref.sh")
D65 = (0.95047, 1.0, 1.0883)      # ScielabProcessor.java:20 (values only select the kernel arguments)
D50 = (0.966797, 1.0, 0.825188)   # :21
_P = C.c_void_p
EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_double))
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build(force: bool = False) -> bool:
    """Compiles oracle/_ref when the reference tree is present (this container); on the GPU box the prebuilt
    .so that travelled with the snapshot is used.  Returns available()."""
    srcs = [os.path.join(_HERE, "ref_build", f) for f in os.listdir(os.path.join(_HERE, "ref_build"))]
    stale = force or not available() or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale and os.path.isdir(os.environ.get("HQ_REFERENCE", "/root/reference")):
        subprocess.run(["bash", BUILD_SCRIPT], check=True)
    return available()


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libhq_ref.so is not built (needs /root/reference; run oracle/ref_build/build_ref.sh)")
        L = C.CDLL(LIB_PATH)
        i, f, d = C.c_int, C.c_float, C.c_double
        L.refcl_RGB2XYZ.argtypes = [_P, _P, _P, _P, i, i]
        L.refcl_XYZ2Opp.argtypes = [_P, _P, i, i]
        L.refcl_Opp2LAB.argtypes = [_P, f, f, f, _P, i, i]
        L.refcl_quantize.argtypes = [_P, _P, i, _P, _P, i, i]
        L.refcl_quantizeAndConvertToOpp.argtypes = [_P, _P, i, _P, _P, i, i]
        L.refcl_CIEDE.argtypes = [_P, _P, _P, i, i]
        L.refcl_convolve4Channels.argtypes = [_P, _P, i, i, i, i, _P, i, i]
        L.refcl_convolve1Channel.argtypes = [_P, _P, i, i, i, i, _P, i, i]
        L.refcl_computeScielabKernelsTemp.argtypes = [_P, _P, _P, _P, i, i, i, _P, _P, _P, i, i]
        L.refcl_computeScielabKernelsEnd.argtypes = [_P, _P, _P, _P, _P, _P, i, i, i, _P, i, i]
        L.refcl_builtin_range.argtypes = [i, C.c_uint, C.c_uint, _P]
        L.refj_sRGBtoLab.argtypes = [_P, _P, _P, i, i, _P, _P, _P]
        L.refj_sRGBtoOpp.argtypes = [_P, _P]
        L.refj_scielab_filters.restype = i; L.refj_scielab_filters.argtypes = [i, d, _P, _P, i]
        L.refj_lab_constant.restype = f; L.refj_lab_constant.argtypes = [i]
        L.refj_seed.argtypes = [C.c_longlong]
        L.refj_draws.restype = C.c_long
        L.refj_nextFloat.restype = f
        L.refj_nextDouble.restype = d
        L.refj_swasa_new.restype = _P; L.refj_swasa_new.argtypes = [i, i, i, f, f, f, f, f, f, f]
        L.refj_swasa_free.argtypes = [_P]
        L.refj_swasa_reset.argtypes = [_P]
        L.refj_swasa_generateRandomColors.argtypes = [_P, i, _P]
        L.refj_swasa_generateNeighboringColors.argtypes = [_P, _P, _P, i, i]
        L.refj_swasa_isAccepted.restype = i; L.refj_swasa_isAccepted.argtypes = [_P, d]
        L.refj_swasa_keepsHisValues.restype = i; L.refj_swasa_keepsHisValues.argtypes = [_P, i]
        L.refj_swasa_acceptanceProbability.restype = d; L.refj_swasa_acceptanceProbability.argtypes = [_P, d]
        L.refj_swasa_maxStepWidth.restype = f; L.refj_swasa_maxStepWidth.argtypes = [_P, i]
        L.refj_swasa_computePenalty.restype = d; L.refj_swasa_computePenalty.argtypes = [_P, _P, i]
        L.refj_swasa_reduceTemperatureIfNecessary.argtypes = [_P, i]
        L.refj_clamp.restype = f; L.refj_clamp.argtypes = [f, f, f]
        L.refj_argmin.restype = i; L.refj_argmin.argtypes = [_P, i]
        L.refj_findBestQuantization.restype = d
        L.refj_findBestQuantization.argtypes = [_P, i, i, EVAL_FN, _P, _P, _P, _P]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def default_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


# ------------------------------------------------------------------ layouts (HybridQuantization.java:279-309)
def unit_planes(rgb_u8) -> np.ndarray:
    """u8 image -> planar float [3][n] in [0,1] (Icy convertToType(FLOAT, rescale), HybridQuantization.java:95 —
    third-party; the build defines it as (float)(c/255.0))"""
    rgb = np.ascontiguousarray(rgb_u8, np.uint8).reshape(-1, 3)
    return np.ascontiguousarray((rgb.astype(np.float64) / 255.0).astype(np.float32).T)


def makeinline(planes) -> np.ndarray:
    """[3][n] -> [n][4] = R,G,B,0 (HybridQuantization.makeinline :279-291)"""
    planes = _f32(planes)
    out = np.zeros((planes.shape[1], 4), np.float32)
    out[:, :3] = planes.T
    return out


def pack_filters(filters, abs3):
    """updateOpenCLFilters (ImageManipulation.java:800-841).  filters: [7][taps] in the order
    (O1g1,O1g2,O1g3,O2g1,O2g2,O3g1,O3g2) = the reference's filters[channel][gaussian]."""
    filters = _f32(filters); abs3 = _f32(abs3)
    taps = filters.shape[1]
    f4 = [np.zeros((taps, 4), np.float32) for _ in range(3)]
    f4[0][:, 0], f4[0][:, 1], f4[0][:, 2] = filters[0], filters[3], filters[5]   # filters[0..2][0]
    f4[1][:, 0], f4[1][:, 1], f4[1][:, 2] = filters[1], filters[4], filters[6]   # filters[0..2][1]
    f4[2][:, 0] = filters[2]                                                       # filters[0][2]
    abs4 = np.zeros((taps, 4), np.float32); abs4[:, 0] = abs3
    return {"filters4": f4, "filter3": filters[2].copy(), "absfilters4": abs4, "absfilter3": abs3.copy(), "half": (taps * 4) // 8}


# ------------------------------------------------------------------ kernels in the reference's host order
def rgb_to_xyz(planes, threads=None) -> np.ndarray:
    """ImageManipulation.RGBtoXYZ :100-152 -> RGB2XYZ kernel; returns [n][4]"""
    planes = _f32(planes); n = planes.shape[1]
    out = np.zeros((n, 4), np.float32)
    load().refcl_RGB2XYZ(_ptr(planes[0]), _ptr(planes[1]), _ptr(planes[2]), _ptr(out), n, threads or default_threads())
    return out


def xyz_to_scielab(xyz4, packed, w, illuminant=D65, threads=None) -> np.ndarray:
    """ImageManipulation.XYZtoScielab :285-370: XYZ2Opp, (H,V) with filters4[0], (H,V accumulate) with filters4[1],
    (H with filters4[2], V accumulate with absfilters4) on channel x, Opp2LAB.  Returns [n][4]."""
    L = load(); t = threads or default_threads()
    xyz4 = _f32(xyz4); n = xyz4.shape[0]; h = n // w; half = packed["half"]
    opp = np.zeros_like(xyz4); tmp = np.zeros_like(xyz4); conv = np.zeros_like(xyz4); lab = np.zeros_like(xyz4)
    L.refcl_XYZ2Opp(_ptr(xyz4), _ptr(opp), n, t)                                                   # :317-318
    f = packed["filters4"]
    L.refcl_convolve4Channels(_ptr(opp), _ptr(f[0]), half, w, h, 0, _ptr(tmp), n, t)               # :323-324
    L.refcl_convolve4Channels(_ptr(tmp), _ptr(f[0]), half, h, w, 0, _ptr(conv), n, t)              # :326-327
    L.refcl_convolve4Channels(_ptr(opp), _ptr(f[1]), half, w, h, 0, _ptr(tmp), n, t)               # :332-333
    L.refcl_convolve4Channels(_ptr(tmp), _ptr(f[1]), half, h, w, 1, _ptr(conv), n, t)              # :335-336
    L.refcl_convolve1Channel(_ptr(opp), _ptr(f[2]), half, w, h, 0, _ptr(tmp), n, t)                # :341-342
    L.refcl_convolve1Channel(_ptr(tmp), _ptr(packed["absfilters4"]), half, h, w, 1, _ptr(conv), n, t)  # :344-347
    L.refcl_Opp2LAB(_ptr(conv), illuminant[0], illuminant[1], illuminant[2], _ptr(lab), n, t)      # :353-354
    return lab


def srgb_to_scielab(rgb_u8, packed, illuminant=D65, threads=None) -> np.ndarray:
    """ScielabProcessor.sRGBToScielab :374-381"""
    rgb = np.ascontiguousarray(rgb_u8, np.uint8)
    return xyz_to_scielab(rgb_to_xyz(unit_planes(rgb), threads), packed, rgb.shape[1], illuminant, threads)


def quantize(rgb4, colors, threads=None):
    """ImageManipulation.quantize :770-798 -> quantize kernel.  Returns (out [n][4], used [K])"""
    rgb4 = _f32(rgb4); colors = _f32(colors); n = rgb4.shape[0]; K = colors.shape[0]
    out = np.zeros_like(rgb4); used = np.zeros(K, np.int32)
    load().refcl_quantize(_ptr(rgb4), _ptr(colors), K, _ptr(used), _ptr(out), n, threads or default_threads())
    return out, used


def average_array(err: np.ndarray, depth: int) -> float:
    """averageArray/sumArray (ImageManipulation.java:736-768): double sums over a binary split of depth `depth`
    (the reference derives depth from availableProcessors(), :738), divided by the length"""
    def rec(lo, hi, d):
        if hi <= lo:
            return 0.0
        if d <= 0:
            if hi - lo > 200000:   # large leaves: a float64 numpy sum (pairwise) instead of the sequential loop; both are double
                return float(err[lo:hi].sum(dtype=np.float64))   # sums of floats and agree to ~1e-16 relative
            s = 0.0
            for v in err[lo:hi].astype(np.float64).tolist():   # sequential, left to right (:748-751)
                s += v
            return s
        mid = (lo + hi) // 2
        return rec(lo, mid, d - 1) + rec(mid, hi, d - 1)
    return rec(0, err.shape[0], depth) / err.shape[0]


def eval_population(rgb4, scielab4, w, packed, palettes, swasa=None, illuminant=D65, threads=None, depth=0, details=False):
    """computeQuantizationErrorPopulation (ImageManipulation.java:620-727): for each candidate the kernels
    quantizeAndConvertToOpp, computeScielabKernelsTemp(w,h), computeScielabKernelsEnd(h,w), Opp2LAB, CIEDE in that
    order (:644-665), then averageArray + computePenalty on the host (:712).  palettes [P][K][4]."""
    L = load(); t = threads or default_threads()
    rgb4 = _f32(rgb4); scielab4 = _f32(scielab4); palettes = _f32(palettes)
    n = rgb4.shape[0]; h = n // w; P, K, _ = palettes.shape; half = packed["half"]
    f = packed["filters4"]
    opp = np.zeros((n, 4), np.float32); t1 = np.zeros((n, 4), np.float32); t2 = np.zeros((n, 4), np.float32)
    t3 = np.zeros(n, np.float32); conv = np.zeros((n, 4), np.float32); lab = np.zeros((n, 4), np.float32)
    costs = np.zeros(P, np.float64); out = []
    for i in range(P):
        used = np.zeros(K, np.int32); err = np.zeros(n, np.float32)
        L.refcl_quantizeAndConvertToOpp(_ptr(rgb4), _ptr(palettes[i]), K, _ptr(used), _ptr(opp), n, t)
        L.refcl_computeScielabKernelsTemp(_ptr(opp), _ptr(f[0]), _ptr(f[1]), _ptr(packed["filter3"]), half, w, h, _ptr(t1), _ptr(t2), _ptr(t3), n, t)
        L.refcl_computeScielabKernelsEnd(_ptr(t1), _ptr(t2), _ptr(t3), _ptr(f[0]), _ptr(f[1]), _ptr(packed["absfilter3"]), half, h, w, _ptr(conv), n, t)
        L.refcl_Opp2LAB(_ptr(conv), illuminant[0], illuminant[1], illuminant[2], _ptr(lab), n, t)
        L.refcl_CIEDE(_ptr(scielab4), _ptr(lab), _ptr(err), n, t)
        pen = L.refj_swasa_computePenalty(swasa, _ptr(used), K) if swasa else 0.0
        costs[i] = average_array(err, depth) + pen
        if details:
            out.append({"used": used, "err": err, "opp": opp.copy(), "lab": lab.copy()})
    return (costs, out) if details else costs


# ------------------------------------------------------------------ Java pieces
def srgb_to_lab_java(planes, d50=False):
    """ScielabProcessor.sRGBtoLab :422-438 = OpptoLab(sRGBtoOpp(px)); planes [3][n] -> [3][n]"""
    planes = _f32(planes); n = planes.shape[1]
    out = np.zeros((3, n), np.float32)
    load().refj_sRGBtoLab(_ptr(planes[0]), _ptr(planes[1]), _ptr(planes[2]), n, int(d50), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]))
    return out


def scielab_filters(dpi=72, viewing_distance=45.0, max_taps=4096):
    """(filters [7][taps], abs3 [taps]) from the ScielabProcessor constructor (ScielabProcessor.java:78-178)"""
    f = np.zeros(7 * max_taps, np.float32); a = np.zeros(max_taps, np.float32)
    t = load().refj_scielab_filters(dpi, float(viewing_distance), _ptr(f), _ptr(a), max_taps)
    if t < 0:
        raise ValueError("too many taps")
    return f[:7 * t].reshape(7, t).copy(), a[:t].copy()


class Swasa:
    """SWASA.java compiled as it stands; the RNG behind icy.util.Random is the pinned java.util.Random LCG"""

    def __init__(self, population=4, imax=5000, iTc=20, delta=2.0, conv_delay=0.75, conv_spread=0.15, t0=20.0, alpha=0.9, s0=100.0, beta=5.3):
        # defaults = the plugin's (HybridQuantization.java:196-224)
        self.h = load().refj_swasa_new(population, imax, iTc, delta, conv_delay, conv_spread, t0, alpha, s0, beta)
        self.population, self.imax = population, imax

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.refj_swasa_free(self.h); self.h = None

    def generateRandomColors(self, K):
        out = np.zeros((K, 4), np.float32); load().refj_swasa_generateRandomColors(self.h, K, _ptr(out)); return out

    def generateNeighboringColors(self, colors, iteration):
        colors = _f32(colors); out = np.zeros_like(colors)
        load().refj_swasa_generateNeighboringColors(self.h, _ptr(colors), _ptr(out), colors.shape[0], iteration); return out

    def maxStepWidth(self, i):
        return load().refj_swasa_maxStepWidth(self.h, i)

    def isAccepted(self, dE):
        return bool(load().refj_swasa_isAccepted(self.h, dE))

    def keepsHisValues(self, it):
        return bool(load().refj_swasa_keepsHisValues(self.h, it))

    def reduceTemperatureIfNecessary(self, it):
        load().refj_swasa_reduceTemperatureIfNecessary(self.h, it)

    def acceptanceProbability(self, dE):
        return load().refj_swasa_acceptanceProbability(self.h, dE)

    def computePenalty(self, used):
        used = np.ascontiguousarray(used, np.int32); return load().refj_swasa_computePenalty(self.h, _ptr(used), used.shape[0])

    def reset(self):
        load().refj_swasa_reset(self.h)


def seed(s: int) -> None:
    load().refj_seed(s)


def find_best_quantization(swasa: Swasa, K: int, evaluate, convergence=True, trace=False):
    """findBestQuantization's annealing (ImageManipulation.java:385, :413-419, :490-545, compiled from the
    reference) around `evaluate(palettes [P][K][4]) -> costs [P]`.  Returns (best [K][4], bestError, trace)."""
    tr = np.zeros((swasa.imax + 1) * swasa.population, np.float64) if trace else None

    def cb(_user, pal, P, Kc, costs):
        a = np.ctypeslib.as_array(pal, shape=(P, Kc, 4)).copy()
        c = np.asarray(evaluate(a), np.float64)
        for j in range(P):
            costs[j] = c[j]

    fn = EVAL_FN(cb)
    best = np.zeros((K, 4), np.float32)
    err = load().refj_findBestQuantization(swasa.h, K, int(convergence), fn, None, _ptr(best), _ptr(tr), None)
    return best, err, (tr.reshape(swasa.imax + 1, swasa.population) if trace else None)


def reference_plugin_search(rgb_u8, K, swasa: Swasa, filters, abs3, illuminant=D65, convergence=True, trace=False, threads=None, depth=0):
    """The plugin's quantization path end to end on the CPU from the reference's own code: sRGBToScielab of the
    original (ScielabProcessor.java:374-381), then bestColors -> findBestQuantization with the OpenCL candidate
    chain (HybridQuantization.java:100-107)."""
    rgb = np.ascontiguousarray(rgb_u8, np.uint8); w = rgb.shape[1]
    packed = pack_filters(filters, abs3)
    scielab4 = srgb_to_scielab(rgb, packed, illuminant, threads)
    rgb4 = makeinline(unit_planes(rgb))
    ev = lambda pal: eval_population(rgb4, scielab4, w, packed, pal, swasa.h, illuminant, threads, depth)
    return find_best_quantization(swasa, K, ev, convergence, trace)


# ---------------------------------------------------------------- the CIEDE kernel built with -DCIE94 (scope row f4)
LIB94_PATH = os.path.join(_HERE, "_ref", "libhq_ref94.so")
_lib94 = None


def ciede94(lab1: np.ndarray, lab2: np.ndarray) -> np.ndarray:
    """The reference's CIEDE kernel compiled with -DCIE94 (cl:217-226) on [n, 3] Lab arrays (padded to float4 as the kernel reads them)"""
    global _lib94
    if _lib94 is None:
        if not os.path.exists(LIB94_PATH):
            raise RuntimeError("oracle/_ref/libhq_ref94.so is not built (run oracle/ref_build/build_ref.sh)")
        _lib94 = C.CDLL(LIB94_PATH)
        _lib94.refcl94_CIEDE.argtypes = [_P, _P, _P, C.c_int]
    a = np.zeros((len(lab1), 4), np.float32); a[:, :3] = lab1
    b = np.zeros((len(lab2), 4), np.float32); b[:, :3] = lab2
    out = np.empty(len(a), np.float32)
    _lib94.refcl94_CIEDE(_ptr(a), _ptr(b), _ptr(out), len(a))
    return out
