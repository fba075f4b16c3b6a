"""hybridquantization_b200 — B200-native (sm_100a) hot path of the HybridQuantization plugin.

RGB -> CIELAB, nearest-palette assignment, per-colour counts / Lab sums / total-error
reduction and batched SWASA candidate scoring, behind the C ABI of include/hq_b200.h.
The CUDA library is required: importing `plugin` objects that touch it raises if
libhq_b200.so has not been built, and creating a backend raises if there is no GPU.
"""
from ._lib import (COST_LAB, COST_SCIELAB, EVAL_FORCE_CHUNKED, EVAL_FORCE_DIRECT, EVAL_FORCE_PREFILTER, EVAL_PRUNE, EVAL_ALLREDUCE, EVAL_SUMS, LIB_PATH, PRUNE_AUTO, PRUNE_OFF, PRUNE_ON, MAX_COLORS, SPACE_LAB, SPACE_SRGB,
                   WHITEPOINT_D50, WHITEPOINT_D65, HqError)
from .plugin import HybridQuantization, ImageManipulation, JavaRandom, ScielabProcessor, SWASA

__all__ = ["HybridQuantization", "ImageManipulation", "JavaRandom", "ScielabProcessor", "SWASA", "HqError",
           "SPACE_LAB", "SPACE_SRGB", "COST_LAB", "COST_SCIELAB", "WHITEPOINT_D65", "WHITEPOINT_D50", "EVAL_SUMS", "EVAL_FORCE_DIRECT",
           "EVAL_FORCE_CHUNKED", "EVAL_FORCE_PREFILTER", "EVAL_PRUNE", "EVAL_ALLREDUCE", "PRUNE_OFF", "PRUNE_AUTO", "PRUNE_ON", "MAX_COLORS", "LIB_PATH"]
