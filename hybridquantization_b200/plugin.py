"""Python mirror of the reference's plugin interface for the hot path, over the C ABI.

Class and method names follow the reference so that tests read like tests of the plugin:
  JavaRandom          icy.util.Random / java.util.Random      (SWASA.java:46-48)
  SWASA               SWASA.java
  ImageManipulation   ImageManipulation.java (the backend the CUDA library replaces)
  ScielabProcessor    ScielabProcessor.java  (white point + bestColors façade)
  HybridQuantization  HybridQuantization.java (parameter surface :185-257, quantization() :93-137)
Everything numeric happens in libhq_b200.so; this file only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (COST_LAB, COST_SCIELAB, EVAL_FORCE_CHUNKED, EVAL_FORCE_DIRECT, EVAL_FORCE_PREFILTER, EVAL_SUMS, SPACE_LAB, SPACE_SRGB, WHITEPOINT_D50,
                   WHITEPOINT_D65, HqError, JavaRandomState, SwasaParams)

__all__ = ["COST_LAB", "COST_SCIELAB", "JavaRandom", "SWASA", "ImageManipulation", "ScielabProcessor", "HybridQuantization", "HqError",
           "SPACE_LAB", "SPACE_SRGB", "WHITEPOINT_D65", "WHITEPOINT_D50"]


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class JavaRandom:
    """java.util.Random LCG with an explicit seed (the reference's generator is unseeded)."""

    def __init__(self, seed: int = 0):
        self._s = JavaRandomState()
        _lib.load().hq_java_random_seed(C.byref(self._s), seed)

    def next(self, bits: int) -> int:
        return _lib.load().hq_java_random_next(C.byref(self._s), bits)

    def nextInt(self) -> int:
        return self.next(32)

    def nextFloat(self) -> float:
        return _lib.load().hq_java_random_next_float(C.byref(self._s))

    def nextDouble(self) -> float:
        return _lib.load().hq_java_random_next_double(C.byref(self._s))


class SWASA:
    """SWASA.java: schedule parameters; the accept/reject loop itself runs in the library's
    host driver (hq_find_best_quantization), which uses the same C++ class."""

    def __init__(self, population=4, imax=5000, iTc=20, delta=2.0, convDelay=0.75, convSpread=0.15, t0=20.0,
                 alpha=0.9, s0=100.0, beta=5.3, seed=77760, convergence=True, space=SPACE_LAB, costModel=COST_LAB):
        p = SwasaParams()
        _lib.load().hq_swasa_default_params(C.byref(p))
        p.population, p.imax, p.iTc, p.delta = population, imax, iTc, delta
        p.convergence, p.conv_delay, p.conv_spread = int(bool(convergence)), convDelay, convSpread
        p.t0, p.alpha, p.s0, p.beta = t0, alpha, s0, beta
        p.space, p.seed, p.cost_model = space, seed, costModel
        self.params = p
        self.random = JavaRandom(seed)

    def getImax(self) -> int:
        return self.params.imax

    def getPopulationSize(self) -> int:
        return self.params.population

    def generateRandomColors(self, numberOfColors: int) -> np.ndarray:
        out = np.empty((numberOfColors, 4), np.float32)
        _lib.load().hq_swasa_generate_random_colors(C.byref(self.random._s), numberOfColors, _ptr(out))
        return out

    def maxStepWidth(self, i: int) -> float:
        return _lib.load().hq_swasa_max_step_width(C.byref(self.params), i)

    def generateNeighboringColors(self, colors: np.ndarray, iteration: int) -> np.ndarray:
        colors = np.ascontiguousarray(colors, np.float32)
        out = np.empty_like(colors)
        _lib.load().hq_swasa_generate_neighboring_colors(C.byref(self.params), C.byref(self.random._s), _ptr(colors),
                                                         _ptr(out), colors.shape[0], iteration)
        return out

    def computePenalty(self, counts: np.ndarray) -> float:
        return float(np.count_nonzero(np.asarray(counts) == 0)) * float(np.float32(self.params.delta))


class ImageManipulation:
    """The compute backend (ImageManipulation.java) on one GPU.  Creation raises HqError when
    no usable device exists — the reference's silent zero-output mode (:79-92) is gone."""

    def __init__(self, deltaEType: str = "CIE76", verbose: bool = False, convergence: bool = True, device=0):
        """device: a CUDA device index, or a list of them -> ONE context over several GPUs of this process (hq_create_multi:
        the image rows are split inside the library, every evaluation ends in one grouped NCCL all-reduce)."""
        if deltaEType not in ("CIE76", "CIE94", "CIEDE2000"):
            raise ValueError("deltaEType must be one of ImageManipulation.deltaETypes (CIE76, CIE94, CIEDE2000)")
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        self.verbose, self.convergence = verbose, convergence
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*device)
            rc = self._lib.hq_create_multi(devs, len(device), C.byref(self._ctx))
        else:
            rc = self._lib.hq_create(device, C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.hq_last_error(None)
            self._ctx = C.c_void_p()
            raise HqError(rc, msg.decode() if msg else "")
        self._cb = None
        self.shape = None
        self._local_pixels = 0
        if deltaEType != "CIE76":   # :63 -D<type>; CIEDE2000 (an empty stub in the reference, cl:227-229) is refused by the library
            try:
                self.setDeltaE({"CIE94": _lib.DELTAE_CIE94, "CIEDE2000": _lib.DELTAE_CIEDE2000}[deltaEType])
            except HqError:
                self.close()
                raise

    def setDeltaE(self, type_: int) -> None:
        _lib.check(self._ctx, self._lib.hq_set_delta_e(self._ctx, type_))

    # -- lifetime
    def getCudaAvailable(self) -> bool:
        return bool(self._ctx)

    def close(self) -> None:
        if self._ctx:
            self._lib.hq_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def deviceInfo(self) -> dict:
        sm, clk = C.c_int(), C.c_int()
        name = C.create_string_buffer(128)
        _lib.check(self._ctx, self._lib.hq_device_info(self._ctx, C.byref(sm), C.byref(clk), name, 128))
        return {"name": name.value.decode(), "sm_count": sm.value, "sm_clock_khz": clk.value}

    # -- native NCCL (one process per GPU): the library owns the communicator, the host only ships the id bytes
    @staticmethod
    def commUniqueId() -> bytes:
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        rc = _lib.load().hq_comm_get_unique_id(buf)
        if rc != 0:
            msg = _lib.load().hq_last_error(None)
            raise HqError(rc, msg.decode() if msg else "")
        return buf.raw

    def commInitRank(self, unique_id: bytes, nranks: int, rank: int) -> None:
        if len(unique_id) != _lib.COMM_ID_BYTES:
            raise ValueError("unique_id must be the %d bytes of commUniqueId()" % _lib.COMM_ID_BYTES)
        _lib.check(self._ctx, self._lib.hq_comm_init_rank(self._ctx, C.c_char_p(unique_id), nranks, rank))

    def commAllreduce(self, d_words_ptr: int, n_words: int, stream: int = 0) -> None:
        _lib.check(self._ctx, self._lib.hq_comm_allreduce(self._ctx, d_words_ptr, n_words, stream or None))

    def commInfo(self) -> dict:
        r, s_, v = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self._ctx, self._lib.hq_comm_info(self._ctx, C.byref(r), C.byref(s_), C.byref(v)))
        return {"rank": r.value, "size": s_.value, "nccl_version": v.value, "devices": int(self._lib.hq_multi_device_count(self._ctx)),
                "peer_exchange": bool(self._lib.hq_comm_peers_open(self._ctx))}

    # -- the exchange of small payloads over NVLink peer memory (hq_b200.h): mailbox handle out, every rank's handles in
    def commPeerHandle(self) -> bytes:
        buf = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        _lib.check(self._ctx, self._lib.hq_comm_peer_handle(self._ctx, buf))
        return buf.raw

    def commOpenPeers(self, handles: list, rank: int) -> None:
        if any(len(h) != _lib.PEER_HANDLE_BYTES for h in handles):
            raise ValueError("every handle must be the %d bytes of commPeerHandle()" % _lib.PEER_HANDLE_BYTES)
        blob = b"".join(handles)
        _lib.check(self._ctx, self._lib.hq_comm_open_peers(self._ctx, C.c_char_p(blob), len(handles), rank))

    def commClosePeers(self) -> None:
        self._lib.hq_comm_close_peers(self._ctx)

    def commPeersOpen(self) -> bool:
        return bool(self._lib.hq_comm_peers_open(self._ctx))

    # -- measurement hooks
    def setProfiling(self, enabled: bool) -> None:
        _lib.check(self._ctx, self._lib.hq_set_profiling(self._ctx, int(enabled)))

    def lastAssignMs(self) -> float:
        ms = C.c_float()
        _lib.check(self._ctx, self._lib.hq_last_assign_ms(self._ctx, C.byref(ms)))
        return ms.value

    def lastRgbToLabMs(self) -> float:
        ms = C.c_float()
        _lib.check(self._ctx, self._lib.hq_last_rgb_to_lab_ms(self._ctx, C.byref(ms)))
        return ms.value

    def lastScielabStageMs(self):
        """(ms, candidates) of the filter-stage launches of the last sub-batch of the latest evalPalettesScielab (profiling on)."""
        ms, nb = C.c_float(), C.c_int()
        _lib.check(self._ctx, self._lib.hq_last_scielab_stage_ms(self._ctx, C.byref(ms), C.byref(nb)))
        return ms.value, nb.value

    def measureFp32Peak(self) -> dict:
        a, b = C.c_double(), C.c_double()
        _lib.check(self._ctx, self._lib.hq_measure_fp32_peak(self._ctx, C.byref(a), C.byref(b)))
        return {"ffma_tflops": a.value, "ffma2_tflops": b.value}

    # -- image
    def setImage(self, rgb: np.ndarray, whitepoint: int = WHITEPOINT_D65) -> None:
        """rgb: uint8 [rows, width, 3] (this rank's rows)."""
        rgb = np.ascontiguousarray(rgb, np.uint8)
        if rgb.ndim != 3 or rgb.shape[2] != 3:
            raise ValueError("Please open an image with 3 or more channels")  # HybridQuantization.java:68-69
        self.shape = rgb.shape[:2]
        self._local_pixels = rgb.shape[0] * rgb.shape[1]
        _lib.check(self._ctx, self._lib.hq_set_image_u8(self._ctx, _ptr(rgb), rgb.shape[1], rgb.shape[0], whitepoint))

    def setImageSharded(self, rgb_local: np.ndarray, halo_top: int, halo_bottom: int, global_row0: int, global_rows: int,
                        whitepoint: int = WHITEPOINT_D65) -> None:
        """rgb_local: uint8 [halo_top + own_rows + halo_bottom, width, 3]; see hq_set_image_u8_sharded."""
        rgb = np.ascontiguousarray(rgb_local, np.uint8)
        own = rgb.shape[0] - halo_top - halo_bottom
        self.shape = (own, rgb.shape[1])
        self._local_pixels = rgb.shape[0] * rgb.shape[1]
        _lib.check(self._ctx, self._lib.hq_set_image_u8_sharded(self._ctx, _ptr(rgb), rgb.shape[1], own, halo_top, halo_bottom,
                                                                  global_row0, global_rows, whitepoint))

    def setImageFloat(self, planes: np.ndarray, whitepoint: int = WHITEPOINT_D65, halo_top: int = 0, halo_bottom: int = 0,
                      global_row0: int = 0, global_rows: int | None = None) -> None:
        """planes: float32 [3, rows, width] in [0,1] — `im.getDataXYCAsFloat()` (HybridQuantization.java:95-98), any Icy
        data type after the rescaling conversion.  With halos: rows = halo_top + own rows + halo_bottom."""
        p = np.ascontiguousarray(planes, np.float32)
        if p.ndim != 3 or p.shape[0] < 3:
            raise ValueError("Please open an image with 3 or more channels")  # HybridQuantization.java:68-69
        own = p.shape[1] - halo_top - halo_bottom
        self.shape = (own, p.shape[2])
        self._local_pixels = p.shape[1] * p.shape[2]
        _lib.check(self._ctx, self._lib.hq_set_image_f32_planar_sharded(
            self._ctx, _ptr(p[0]), _ptr(p[1]), _ptr(p[2]), p.shape[2], own, halo_top, halo_bottom, global_row0,
            own if global_rows is None else global_rows, whitepoint))

    def setImageDevice(self, d_rgb_ptr: int, width: int, rows: int, whitepoint: int = WHITEPOINT_D65, stream: int = 0):
        self.shape = (rows, width)
        self._local_pixels = rows * width
        _lib.check(self._ctx, self._lib.hq_set_image_u8_device(self._ctx, d_rgb_ptr, width, rows, whitepoint, stream or None))

    def pixels(self) -> int:
        return int(self._lib.hq_image_pixels(self._ctx))

    def labImage(self) -> np.ndarray:
        out = np.empty((3, self.pixels()), np.float32)
        _lib.check(self._ctx, self._lib.hq_get_lab(self._ctx, _ptr(out)))
        return out

    # -- evaluation
    def evalPalettes(self, palettes: np.ndarray, space: int = SPACE_LAB, sums: bool = False, flags: int = 0) -> dict:
        """palettes float32 [B, K, 4] (R,G,B,0).  Returns exact integer reductions."""
        palettes = np.ascontiguousarray(palettes, np.float32)
        if palettes.ndim == 2:
            palettes = palettes[None]
        B, K, four = palettes.shape
        if four != 4:
            raise ValueError("palettes must be [B, K, 4]")
        err = np.empty(B, np.int64)
        counts = np.empty((B, K), np.uint64)
        sums_fx = np.empty((B, K, 3), np.int64) if sums else None
        fl = flags | (EVAL_SUMS if sums else 0)
        _lib.check(self._ctx, self._lib.hq_eval_palettes(self._ctx, _ptr(palettes), B, K, space, fl, _ptr(err), _ptr(counts), _ptr(sums_fx)))
        return {"err_fx": err, "counts": counts, "sums_fx": sums_fx}

    def evalPalettesDevice(self, d_palettes_ptr: int, B: int, K: int, d_results_ptr: int, space: int = SPACE_LAB,
                           flags: int = 0, stream: int = 0) -> None:
        _lib.check(self._ctx, self._lib.hq_eval_palettes_device(self._ctx, d_palettes_ptr, B, K, space, flags, d_results_ptr, stream or None))

    def resultWords(self, K: int, flags: int = 0) -> int:
        return self._lib.hq_result_words(K, flags)

    def cost(self, err_fx: int, counts: np.ndarray, n_total: int, delta: float) -> float:
        counts = np.ascontiguousarray(counts, np.uint64)
        return self._lib.hq_cost(int(err_fx), _ptr(counts), counts.shape[0], n_total, delta)

    def computeQuantizationErrorPopulation(self, colors: np.ndarray, swasa: SWASA, n_total: int = 0, space: int = SPACE_LAB) -> np.ndarray:
        """ImageManipulation.java:620-727: one cost per candidate palette."""
        r = self.evalPalettes(colors, space)
        n_total = n_total or self.pixels()
        return np.array([self.cost(r["err_fx"][i], r["counts"][i], n_total, swasa.params.delta) for i in range(len(r["err_fx"]))])

    # -- S-CIELAB stage (the plugin's real cost; scope row "next 1")
    def scielabConfigure(self, dpi: int = 72, viewingDistance: float = 45.0) -> None:
        _lib.check(self._ctx, self._lib.hq_scielab_configure(self._ctx, dpi, viewingDistance))

    def scielabSetFilters(self, filters7: np.ndarray, abs3: np.ndarray) -> None:
        filters7 = np.ascontiguousarray(filters7, np.float32); abs3 = np.ascontiguousarray(abs3, np.float32)
        _lib.check(self._ctx, self._lib.hq_scielab_set_filters(self._ctx, _ptr(filters7), _ptr(abs3), filters7.shape[1]))

    def scielabForceGeneric(self, mode) -> None:
        """test hook: 0 default (fused candidate kernel), 1 generic any-tap kernels, 2 round 1's two-kernel 21-tap candidate path"""
        _lib.check(self._ctx, self._lib.hq_scielab_force_generic(self._ctx, int(mode)))

    def scielabFilters(self):
        taps = C.c_int(0)
        _lib.check(self._ctx, self._lib.hq_scielab_get_filters(self._ctx, None, None, C.byref(taps)))
        f = np.empty((7, taps.value), np.float32); a = np.empty(taps.value, np.float32)
        _lib.check(self._ctx, self._lib.hq_scielab_get_filters(self._ctx, _ptr(f), _ptr(a), C.byref(taps)))
        return f, a

    def scielabImage(self) -> np.ndarray:
        """sRGBToScielab(original) (ScielabProcessor.java:374-381): planes [3, n]"""
        out = np.empty((3, self.pixels()), np.float32)
        _lib.check(self._ctx, self._lib.hq_scielab_get_image(self._ctx, _ptr(out)))
        return out

    def evalPalettesScielab(self, palettes: np.ndarray, space: int = SPACE_SRGB) -> dict:
        palettes = np.ascontiguousarray(palettes, np.float32)
        if palettes.ndim == 2:
            palettes = palettes[None]
        B, K, four = palettes.shape
        if four != 4:
            raise ValueError("palettes must be [B, K, 4]")
        err = np.empty(B, np.int64); counts = np.empty((B, K), np.uint64)
        _lib.check(self._ctx, self._lib.hq_eval_palettes_scielab(self._ctx, _ptr(palettes), B, K, space, _ptr(err), _ptr(counts)))
        return {"err_fx": err, "counts": counts}

    # -- the reference class's own entries on its interleaved float4 layouts (what CudaImageManipulation.java calls)
    def RGBtoXYZ(self, R: np.ndarray, G: np.ndarray, B: np.ndarray) -> np.ndarray:
        """ImageManipulation.RGBtoXYZ (:100-152): planar sRGB floats in [0,1] -> XYZ float32 [n, 4]"""
        R, G, B = (np.ascontiguousarray(a, np.float32).ravel() for a in (R, G, B))
        out = np.empty((R.size, 4), np.float32)
        _lib.check(self._ctx, self._lib.hq_rgb_to_xyz(self._ctx, _ptr(R), _ptr(G), _ptr(B), R.size, _ptr(out)))
        return out

    def XYZtoScielab(self, XYZ: np.ndarray, w: int, illuminant) -> np.ndarray:
        """ImageManipulation.XYZtoScielab (:285-370) with the context's filter bank: XYZ [n, 4] -> S-CIELAB [n, 4]"""
        XYZ = np.ascontiguousarray(XYZ, np.float32).reshape(-1, 4)
        ill = np.ascontiguousarray(illuminant, np.float32)
        out = np.empty_like(XYZ)
        _lib.check(self._ctx, self._lib.hq_xyz_to_scielab(self._ctx, _ptr(XYZ), w, XYZ.shape[0] // w, _ptr(ill), _ptr(out)))
        return out

    def scielabSetImage(self, lab4: np.ndarray) -> None:
        """findBestQuantization's inlineScielabOriginal argument (:383): the caller's S-CIELAB image [n, 4] becomes the target"""
        lab4 = np.ascontiguousarray(lab4, np.float32).reshape(-1, 4)
        if lab4.shape[0] != self.pixels():
            raise ValueError("the S-CIELAB image must have one float4 per pixel of the resident image")
        _lib.check(self._ctx, self._lib.hq_scielab_set_image(self._ctx, _ptr(lab4)))

    def computeErrorLab(self, original: np.ndarray, quantized: np.ndarray, errorImage: np.ndarray | None = None) -> float:
        """ImageManipulation.computeError (:858-894) as the reference declares it: two Lab images [n, 4] -> mean dE; errorImage
        [n, 4] (optional) receives ((255 - e)^2) / 255^2 in its first three lanes"""
        a = np.ascontiguousarray(original, np.float32).reshape(-1, 4)
        b = np.ascontiguousarray(quantized, np.float32).reshape(-1, 4)
        if a.shape != b.shape:
            raise ValueError("Mismatching image sizes or not enough channels, abort.")
        if errorImage is not None and (errorImage.dtype != np.float32 or not errorImage.flags.c_contiguous or errorImage.size != a.size):
            raise ValueError("errorImage must be a contiguous float32 array of the images' size")
        mean = C.c_double()
        _lib.check(self._ctx, self._lib.hq_delta_e_images(self._ctx, _ptr(a), _ptr(b), a.shape[0], _ptr(errorImage), C.byref(mean)))
        return mean.value

    def computeError(self, quantized_rgb: np.ndarray) -> dict:
        """Error-image mode (ImageManipulation.computeError :858-894): mean dE between S-CIELAB(original) and
        S-CIELAB(quantized) and the ((255-dE)^2)/255^2 map."""
        q = np.ascontiguousarray(quantized_rgb, np.uint8)
        if q.size != self._local_pixels * 3:  # a shard passes its own rows plus the same halo rows as the original
            raise ValueError("Mismatching image sizes or not enough channels, abort.")  # HybridQuantization.java:81
        emap = np.empty(self.pixels(), np.float32); e8 = np.empty(self.pixels(), np.uint8)
        mean = C.c_double()
        _lib.check(self._ctx, self._lib.hq_error_image(self._ctx, _ptr(q), _ptr(emap), _ptr(e8), C.byref(mean)))
        if self.shape is not None:
            emap, e8 = emap.reshape(self.shape), e8.reshape(self.shape)
        return {"deltaE": mean.value, "errorImage": emap, "errorImageU8": e8}

    def setAllreduce(self, fn) -> None:
        """fn(d_words_ptr: int, n_words: int, stream: int) -> 0 on success; None removes the hook."""
        if fn is None:
            self._cb = None
            _lib.check(self._ctx, self._lib.hq_set_allreduce(self._ctx, _lib.ALLREDUCE_FN(), None))
            return

        def tramp(_user, d_words, n_words, stream):
            try:
                return int(fn(d_words or 0, n_words, stream or 0) or 0)
            except Exception as exc:  # surfaces as HQ_ERR_CALLBACK
                print(f"[hq] all-reduce hook raised: {exc!r}", flush=True)
                return 1

        self._cb = _lib.ALLREDUCE_FN(tramp)
        _lib.check(self._ctx, self._lib.hq_set_allreduce(self._ctx, self._cb, None))

    def findBestQuantization(self, nbOfColors: int, simulatedAnnealing: SWASA, n_total: int = 0, trace: bool = False):
        """ImageManipulation.java:383-591.  Returns (bestColors [K,4], bestError, trace|None, iterations)."""
        p = simulatedAnnealing.params
        p.convergence = int(bool(self.convergence and p.convergence))
        best = np.empty((nbOfColors, 4), np.float32)
        best_err, its = C.c_double(), C.c_int()
        tr = np.empty(((p.imax + 1), p.population), np.float64) if trace else None
        _lib.check(self._ctx, self._lib.hq_find_best_quantization(self._ctx, nbOfColors, C.byref(p), n_total, _ptr(best),
                                                                    C.byref(best_err), _ptr(tr), C.byref(its)))
        return best, best_err.value, tr, its.value

    def setProgress(self, fn) -> None:
        """fn(iteration, max_iterations, best_error) every 10 iterations (the plugin's progress bar, :546-551)."""
        self._progress_cb = _lib.PROGRESS_FN((lambda _u, i, m, e: fn(i, m, e)) if fn else 0)
        _lib.check(self._ctx, self._lib.hq_set_progress(self._ctx, self._progress_cb, None))

    def setPruning(self, mode: int) -> None:
        """PRUNE_OFF | PRUNE_AUTO (default) | PRUNE_ON: whether the search scores populations with the exact pruned kernel"""
        _lib.check(self._ctx, self._lib.hq_set_pruning(self._ctx, mode))

    def searchEvalFlags(self, nbOfColors: int, space: int = 0, costModel: int = 0) -> int:
        """the eval flags the built-in search would use for this image and K (hq_search_eval_flags): EVAL_PRUNE or 0"""
        return int(self._lib.hq_search_eval_flags(self._ctx, nbOfColors, space, costModel))

    def setGraphs(self, enabled: bool) -> None:
        """CUDA-graph replay of repeated identical evalPalettes calls (off by default, see hq_set_graphs)"""
        _lib.check(self._ctx, self._lib.hq_set_graphs(self._ctx, int(bool(enabled))))

    def pruningStats(self) -> dict:
        ch, ms = C.c_uint32(), C.c_double()
        _lib.check(self._ctx, self._lib.hq_pruning_stats(self._ctx, C.byref(ch), C.byref(ms)))
        return {"chunks": ch.value, "mean_survivors": ms.value}

    def requestStop(self) -> None:
        self._lib.hq_request_stop(self._ctx)

    def computeErrorFloat(self, planes: np.ndarray) -> dict:
        """computeError with the second image as float planes [3, rows, width] in [0,1] (HybridQuantization.java:142-143)"""
        p = np.ascontiguousarray(planes, np.float32)
        if p.ndim != 3 or p.shape[0] < 3 or p.shape[1] * p.shape[2] != self._local_pixels:
            raise ValueError("Mismatching image sizes or not enough channels, abort.")  # HybridQuantization.java:81
        emap = np.empty(self.pixels(), np.float32); e8 = np.empty(self.pixels(), np.uint8)
        mean = C.c_double()
        _lib.check(self._ctx, self._lib.hq_error_image_f32_planar(self._ctx, _ptr(p[0]), _ptr(p[1]), _ptr(p[2]), _ptr(emap), _ptr(e8), C.byref(mean)))
        if self.shape is not None:
            emap, e8 = emap.reshape(self.shape), e8.reshape(self.shape)
        return {"deltaE": mean.value, "errorImage": emap, "errorImageU8": e8}

    def quantize(self, colors: np.ndarray, space: int = SPACE_LAB, want_f32: bool = False) -> dict:
        """ImageManipulation.java:770-798 -> packed u8 image, indices, optionally the float RGBA image."""
        colors = np.ascontiguousarray(colors, np.float32)
        n = self.pixels()
        rgb = np.empty((n, 3), np.uint8)
        idx = np.empty(n, np.uint16)
        f32 = np.empty((n, 4), np.float32) if want_f32 else None
        _lib.check(self._ctx, self._lib.hq_quantize(self._ctx, _ptr(colors), colors.shape[0], space, _ptr(rgb), _ptr(f32), _ptr(idx)))
        if self.shape is not None:
            rgb = rgb.reshape(self.shape[0], self.shape[1], 3)
        return {"rgb": rgb, "idx": idx, "f32": f32}


class ScielabProcessor:
    """ScielabProcessor.java: white point, the S-CIELAB filter bank (:66-181) and the bestColors façade."""

    D50, D65 = "D50", "D65"

    @staticmethod
    def buildFilters(dpi: int = 72, viewingDistance: float = 45.0):
        """(Ofilters flattened [7, taps] = O1g1,O1g2,O1g3,O2g1,O2g2,O3g1,O3g2, absOfilters [taps]); host only"""
        lib = _lib.load()
        taps = C.c_int(0)
        if lib.hq_scielab_build_filters(dpi, viewingDistance, None, None, C.byref(taps)) != 0:
            raise ValueError("bad dpi / viewing distance")
        f = np.empty((7, taps.value), np.float32); a = np.empty(taps.value, np.float32)
        lib.hq_scielab_build_filters(dpi, viewingDistance, _ptr(f), _ptr(a), C.byref(taps))
        return f, a

    def __init__(self, dpi: int = 72, viewingDistance: float = 45.0, whitepoint: str = "D65", imageProcessor: ImageManipulation | None = None):
        self.dpi, self.viewingDistance = dpi, viewingDistance
        self.whitepoint = WHITEPOINT_D50 if whitepoint == "D50" else WHITEPOINT_D65
        self.imageProcessing = imageProcessor

    def sRGBToScielab(self, rgb: np.ndarray) -> None:
        """uint8 [rows, width, 3], or the plugin's float planes [3, rows, width] in [0,1] (float[][] sRGBImage, :374)"""
        if np.asarray(rgb).dtype == np.uint8:
            self.imageProcessing.setImage(rgb, self.whitepoint)
        else:
            self.imageProcessing.setImageFloat(rgb, self.whitepoint)

    def bestColors(self, nbOfColors: int, simulatedAnnealing: SWASA, n_total: int = 0):
        return self.imageProcessing.findBestQuantization(nbOfColors, simulatedAnnealing, n_total)

    def close(self) -> None:
        self.imageProcessing.close()


@dataclass
class HybridQuantization:
    """The plugin's parameters (HybridQuantization.java:185-257; names as the EzVar fields) and
    its quantization() entry (:93-137) without the Icy GUI objects."""

    nbOfColors: int = 8
    populationSize: int = 4
    imax: int = 5000
    delta: float = 2.0
    ConvEnable: bool = True
    ConvDelay: float = 0.75
    ConvSpread: float = 0.15
    T0: float = 20.0
    iTc: int = 20
    alpha: float = 0.9
    s0: float = 100.0
    beta: float = 5.3
    dpi: int = 72
    ViewingDistance: float = 45.0
    WhitePoint: str = "D65"
    Verbose: bool = False
    # added: reproducibility, assignment space, device
    seed: int = 77760
    space: int = SPACE_LAB
    costModel: int = COST_LAB  # COST_SCIELAB scores exactly like the reference plugin (with space=SPACE_SRGB)
    device: int = 0
    result: dict = field(default_factory=dict, repr=False)

    def errorImage(self, original: np.ndarray, quantized: np.ndarray) -> dict:
        """HybridQuantization.errorImage (:139-182): S-CIELAB dE image between two images."""
        if original is None or original.size == 0:
            raise ValueError("Please open/select the original image first.")  # :76-77
        if quantized is None or quantized.size == 0:
            raise ValueError("Please open/select the quantized image first.")  # :78-79
        planar = original.dtype != np.uint8  # float planes [3, rows, width] (both converted to FLOAT, :142-143) or u8 [rows, width, 3]
        if original.shape != quantized.shape or original.dtype != quantized.dtype or original.shape[0 if planar else -1] < 3:
            raise ValueError("Mismatching image sizes or not enough channels, abort.")  # :80-81
        imageProcessor = ImageManipulation("CIE76", self.Verbose, False, self.device)  # :145
        try:
            wp = WHITEPOINT_D50 if self.WhitePoint == "D50" else WHITEPOINT_D65
            imageProcessor.setImageFloat(original, wp) if planar else imageProcessor.setImage(original, wp)
            imageProcessor.scielabConfigure(self.dpi, self.ViewingDistance)  # :146
            return imageProcessor.computeErrorFloat(quantized) if planar else imageProcessor.computeError(quantized)  # :153,160
        finally:
            imageProcessor.close()

    def makeSWASA(self) -> SWASA:
        return SWASA(self.populationSize, self.imax, self.iTc, self.delta, self.ConvDelay, self.ConvSpread, self.T0,
                     self.alpha, self.s0, self.beta, seed=self.seed, convergence=self.ConvEnable, space=self.space, costModel=self.costModel)

    def quantization(self, rgb: np.ndarray) -> dict:
        """rgb: uint8 [rows, width, 3], or float32 planes [3, rows, width] in [0,1] (im.getDataXYCAsFloat(), :95-98)"""
        if rgb is None or rgb.size == 0:
            raise ValueError("Please open an image first.")  # :65-67
        imageProcessor = ImageManipulation("CIE76", self.Verbose, self.ConvEnable, self.device)  # :96
        try:
            swasa = self.makeSWASA()  # :97
            scielabProcessor = ScielabProcessor(self.dpi, self.ViewingDistance, self.WhitePoint, imageProcessor)  # :101
            scielabProcessor.sRGBToScielab(rgb)  # :104
            if self.costModel == COST_SCIELAB:
                imageProcessor.scielabConfigure(self.dpi, self.ViewingDistance)  # :101,180
            best, err, _, its = scielabProcessor.bestColors(self.nbOfColors, swasa)  # :107
            q = imageProcessor.quantize(best, self.space)  # :109
            self.result = {"bestColors": best, "bestError": err, "iterations": its, "image": q["rgb"], "idx": q["idx"]}
            return self.result
        finally:
            imageProcessor.close()  # :136
